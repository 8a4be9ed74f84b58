// Fused optimiser tail of the TextBoost step (train_textboost.py:1109-1149 + accelerate GradScaler):
//   grads (already all-reduced, still loss-scaled) -> optional --mixing mask on lora_B rows ->
//   found-inf check + global L2 norm of the LoRA grads -> unscale, 1/world, clip (LoRA only),
//   decoupled AdamW (LoRA lr / embedding lr) -> renormalise the added embedding rows to <= mean_norm
//   -> GradScaler update, step counter, lazy weight-decay scalar of the frozen embedding rows (D8).
// Everything reads its control scalars from device memory: no host synchronisation, graph-capturable.
//
// Flat fp32 layout shared by params / grads / exp_avg / exp_avg_sq:  [ LoRA (n_lora) | added rows (n_rows*D) ]
// Device state vector (fp32[16]):
//   [0] loss_scale  [1] growth_tracker  [2] found_inf  [3] sum g^2 (LoRA, scaled)  [4] step t
//   [5] frozen-row decay c  [6] last clip coefficient  [7] last grad norm (unscaled)  [8] skipped steps
//   [10] fixed-scale flag (0 = GradScaler dynamics, the fp16 policy; 1 = accelerate creates no scaler, bf16)
//   [9] learning-rate multiplier of the step just taken (the --lr_scheduler value, train_textboost.py:911-916)
#include "host_util.h"
#include "sm100.cuh"

namespace tb {

struct AdamCfg {
  long long n_lora, n_total;
  int D, n_rows;
  float lr_lora, lr_emb, beta1, beta2, eps, wd, max_norm, inv_world, mean_norm;
  float growth_factor, backoff_factor;
  int growth_interval;
  int sched_kind;               // TB_LR_* (diffusers.optimization.get_scheduler names)
  float sched_warmup, sched_total;
};

// Multiplier of both learning rates at optimiser step `s` (0-based count of SUCCESSFUL steps so far: accelerate does
// not advance the scheduler when the GradScaler skipped the step).  Restates diffusers.optimization's LambdaLR
// factories with their defaults (num_cycles 0.5 / 1, power 1, lr_end 1e-7) as called at train_textboost.py:911-916;
// accelerate steps the scheduler num_processes times per step with both lengths pre-multiplied by num_processes,
// which leaves these ratio formulas unchanged.
__device__ __forceinline__ float lr_multiplier(const AdamCfg& c, float s, float lr_init) {
  const float warm = c.sched_warmup, total = c.sched_total;
  if (c.sched_kind == 0) return 1.f;
  if (s < warm) return s / fmaxf(1.f, warm);
  if (c.sched_kind == 1) return 1.f;
  const float progress = (s - warm) / fmaxf(1.f, total - warm);
  switch (c.sched_kind) {
    case 2: return fmaxf(0.f, (total - s) / fmaxf(1.f, total - warm));
    case 3: return fmaxf(0.f, 0.5f * (1.f + cosf(3.14159265358979f * progress)));  // num_cycles = 0.5
    case 4: return progress >= 1.f ? 0.f : fmaxf(0.f, 0.5f * (1.f + cosf(3.14159265358979f * fmodf(progress, 1.f))));  // one cycle
    case 5: {
      const float lr_end = 1e-7f;
      if (s > total) return lr_end / lr_init;
      const float pct = 1.f - (s - warm) / (total - warm);
      return ((lr_init - lr_end) * pct + lr_end) / lr_init;
    }
    default: return 1.f;
  }
}

__global__ void optim_mix_mask_kernel(float* __restrict__ g, long long n, int D, int r, int parity) {
  // lora_B blocks are [D][r]; zero rows with (row % 2 == parity)  (train_textboost.py:1119-1126)
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = (i / r) % D;
    if ((row & 1) == parity) g[i] = 0.f;
  }
}

__global__ void optim_reduce_kernel(const float* __restrict__ g, AdamCfg c, float* __restrict__ state) {
  float sq = 0.f;
  int bad = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < c.n_total;
       i += (long long)gridDim.x * blockDim.x) {
    const float v = g[i];
    if (!isfinite(v)) bad = 1;
    if (i < c.n_lora) sq += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  bad = __any_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0) {
    if (sq != 0.f) atomicAdd(&state[3], sq);
    if (bad) state[2] = 1.f;
  }
}

__global__ void optim_update_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                    float* __restrict__ v, AdamCfg c, const float* __restrict__ state) {
  const float found_inf = state[2];
  const float inv_scale = c.inv_world / state[0];
  const float norm = sqrtf(state[3]) * inv_scale;
  const float clip = fminf(1.f, c.max_norm / (norm + 1e-6f));
  const float t = state[4] + 1.f;
  const float bc1 = 1.f - powf(c.beta1, t);
  const float bc2 = 1.f - powf(c.beta2, t);
  const float rsq_bc2 = rsqrtf(bc2);
  const bool finite_norm = isfinite(norm);
  const float mult_lora = lr_multiplier(c, state[4], c.lr_lora), mult_emb = lr_multiplier(c, state[4], c.lr_emb);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < c.n_total;
       i += (long long)gridDim.x * blockDim.x) {
    const float gs = g[i];
    g[i] = 0.f;  // optimizer.zero_grad(): the next step accumulates from zero
    if (found_inf != 0.f || !finite_norm) continue;
    const bool lora = i < c.n_lora;
    const float lr = lora ? c.lr_lora * mult_lora : c.lr_emb * mult_emb;
    const float gr = gs * inv_scale * ((lora && c.max_norm > 0.f) ? clip : 1.f);
    float w = p[i] * (1.f - lr * c.wd);
    const float mi = c.beta1 * m[i] + (1.f - c.beta1) * gr;
    const float vi = c.beta2 * v[i] + (1.f - c.beta2) * gr * gr;
    m[i] = mi;
    v[i] = vi;
    w -= (lr / bc1) * mi / (sqrtf(vi) * rsq_bc2 + c.eps);
    p[i] = w;
  }
}

// one warp per added row: w <- w * min(mean_norm, |w|) / |w|   (train_textboost.py:1138-1149)
// thread 0 of block 0 then advances the scalar state.
__global__ void optim_finish_kernel(float* __restrict__ p, AdamCfg c, float* __restrict__ state,
                                    float* __restrict__ row_norm_mean) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __shared__ float snorm;
  if (threadIdx.x == 0) snorm = 0.f;
  __syncthreads();
  const bool skipped = state[2] != 0.f || !isfinite(state[3]);
  for (int r = warp; r < c.n_rows; r += nw) {
    float* w = p + c.n_lora + (long long)r * c.D;
    float sq = 0.f;
    for (int k = lane; k < c.D; k += 32) sq += w[k] * w[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float nrm = sqrtf(sq);
    const float sc = fminf(c.mean_norm, nrm) / nrm;
    for (int k = lane; k < c.D; k += 32) w[k] *= sc;
    if (lane == 0) atomicAdd(&snorm, nrm);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (row_norm_mean) *row_norm_mean = c.n_rows > 0 ? snorm / c.n_rows : 0.f;
    const float inv_scale = c.inv_world / state[0];
    state[7] = sqrtf(state[3]) * inv_scale;
    state[6] = fminf(1.f, c.max_norm / (state[7] + 1e-6f));
    const bool dynamic_scale = state[10] == 0.f;  // [10] != 0: no GradScaler (bf16 policy), the scale stays put
    if (skipped) {
      if (dynamic_scale) state[0] *= c.backoff_factor;  // GradScaler: halve and skip
      state[1] = 0.f;
      state[8] += 1.f;
    } else {
      state[9] = lr_multiplier(c, state[4], c.lr_lora);
      state[5] *= (1.f - c.lr_emb * lr_multiplier(c, state[4], c.lr_emb) * c.wd);  // frozen rows decay too
      state[4] += 1.f;
      state[1] += 1.f;
      if (state[1] >= (float)c.growth_interval) {
        if (dynamic_scale) state[0] *= c.growth_factor;
        state[1] = 0.f;
      }
    }
    state[2] = 0.f;
    state[3] = 0.f;
  }
}

}  // namespace tb

using namespace tb;

extern "C" int tb_optim_mix_mask(float* grad_lora_b, int64_t n, int D, int r, int parity, void* stream) {
  int rc = tb_check_device();
  if (rc) return rc;
  TB_REQUIRE(grad_lora_b && n % ((int64_t)D * r) == 0 && (parity == 0 || parity == 1), TB_E_ARG,
             "tb_optim_mix_mask: bad args");
  optim_mix_mask_kernel<<<(unsigned)((n + 255) / 256 > 1024 ? 1024 : (n + 255) / 256), 256, 0,
                          (cudaStream_t)stream>>>(grad_lora_b, n, D, r, parity);
  return check_launch("optim_mix_mask_kernel");
}

extern "C" int tb_adamw_fused_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq,
                                   int64_t n_lora, int n_rows, int D, float lr_lora, float lr_emb,
                                   float beta1, float beta2, float eps, float weight_decay,
                                   float max_grad_norm, float inv_world, float mean_norm, int lr_schedule,
                                   float lr_warmup_steps, float lr_total_steps, float* state,
                                   float* row_norm_mean, void* stream) {
  int rc = tb_check_device();
  if (rc) return rc;
  TB_REQUIRE(params && grads && exp_avg && exp_avg_sq && state, TB_E_ARG, "tb_adamw_fused_step: null pointer");
  TB_REQUIRE(lr_schedule >= 0 && lr_schedule <= 5 && lr_warmup_steps >= 0.f, TB_E_ARG,
             "tb_adamw_fused_step: unknown lr_schedule %d", lr_schedule);
  AdamCfg c;
  c.n_lora = n_lora;
  c.n_rows = n_rows;
  c.D = D;
  c.n_total = n_lora + (long long)n_rows * D;
  c.lr_lora = lr_lora;
  c.lr_emb = lr_emb;
  c.beta1 = beta1;
  c.beta2 = beta2;
  c.eps = eps;
  c.wd = weight_decay;
  c.max_norm = max_grad_norm;
  c.inv_world = inv_world;
  c.mean_norm = mean_norm;
  c.growth_factor = 2.f;
  c.backoff_factor = 0.5f;
  c.growth_interval = 2000;
  c.sched_kind = lr_schedule;
  c.sched_warmup = lr_warmup_steps;
  c.sched_total = lr_total_steps;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)((c.n_total + 255) / 256 > 592 ? 592 : (c.n_total + 255) / 256);
  optim_reduce_kernel<<<blocks, 256, 0, st>>>(grads, c, state);
  if ((rc = check_launch("optim_reduce_kernel"))) return rc;
  optim_update_kernel<<<blocks, 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, c, state);
  if ((rc = check_launch("optim_update_kernel"))) return rc;
  optim_finish_kernel<<<1, 256, 0, st>>>(params, c, state, row_norm_mean);
  return check_launch("optim_finish_kernel");
}
