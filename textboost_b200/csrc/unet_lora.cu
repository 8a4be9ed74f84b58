// UNet cross-attention K/V LoRA (--unet_params_to_train crossattn_kv, train_textboost.py:712-721): peft Linear adapters
// on attn2.to_k / attn2.to_v of every transformer block.  The 16 blocks' K | V projections are ONE GEMM over the text
// states here (unet.py: kv_all = ehs x [W_k0 | W_v0 | W_k1 | ...]^T), so the 32 adapters are handled side by side too:
//   A  fp32 [R = n_adapters * r, ctx]   all lora_A stacked (adapter a owns rows a*r .. a*r + r - 1)
//   B  fp32 [KV, r]                     row j = the lora_B row of output column j of the fused projection
//   blk int32 [KV]                      adapter index of output column j;  off int32 [n_adapters + 1] its column range
// forward   Z = ehs A^T (fp32 [M, R]);  kv[m, j] += s * sum_k Z[m, blk[j] r + k] B[j, k]
// backward  dZ[m, a r + k] = s * sum_{j in adapter a} dkv[m, j] B[j, k]
//           dB[j, k] += s * sum_m dkv[m, j] Z[m, blk[j] r + k];  dA[q, c] += sum_m dZ[m, q] ehs[m, c]
//           d_ehs[m, c] += sum_q dZ[m, q] A[q, c]
// M = 616 text rows, R = 128, KV = 24960 on SD-1.5: 0.1 GFLOP of fp32 SIMT work per direction, bound by one pass over
// the [M, KV] 16-bit K/V (30 MB) -- coalesced along KV, the small operands (Z, A, B) stay in L1 / L2.  Masters, the
// down-projection and every sum are fp32; only ehs and kv / dkv are in the build's 16-bit type.
#include "host_util.h"
#include "sm100.cuh"

namespace tb {

constexpr int UL_RMAX = 16;  // rank per adapter

// Z[m, q] = sum_c ehs[m, c] A[q, c]: one CTA per text row (cached in shared memory as fp32), one warp per q
__global__ void __launch_bounds__(256)
unet_lora_down_kernel(const half_t* __restrict__ ehs, const float* __restrict__ A, float* __restrict__ Z, int ctx, int R) {
  extern __shared__ float smf[];  // the text row as fp32
  const int m = blockIdx.x;
  for (int c = threadIdx.x; c < ctx; c += blockDim.x) smf[c] = h2f(ehs[(size_t)m * ctx + c]);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (int q = warp; q < R; q += nwarp) {
    const float* a = A + (size_t)q * ctx;
    float acc = 0.f;
    for (int c = lane; c < ctx; c += 32) acc += smf[c] * a[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) Z[(size_t)m * R + q] = acc;
  }
}

// kv[m, j] += s * Z[m, blk[j] r ..] . B[j, :]: two adjacent columns per thread (one packed load / store)
__global__ void __launch_bounds__(256)
unet_lora_up_kernel(half_t* __restrict__ kv, const float* __restrict__ Z, const float* __restrict__ B,
                    const int* __restrict__ blk, int KV, int R, int r, float s) {
  const int m = blockIdx.y;
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (j >= KV) return;
  half2_t* p = reinterpret_cast<half2_t*>(kv + (size_t)m * KV + j);
  float2 v = h22f2(*p);
  const float* z0 = Z + (size_t)m * R + blk[j] * r;
  const float* z1 = Z + (size_t)m * R + blk[j + 1] * r;
  float a0 = 0.f, a1 = 0.f;
  for (int k = 0; k < r; ++k) {
    a0 += z0[k] * B[(size_t)j * r + k];
    a1 += z1[k] * B[(size_t)(j + 1) * r + k];
  }
  v.x += s * a0;
  v.y += s * a1;
  *p = ff2h2(v.x, v.y);
}

// dZ[m, a r + k] = s * sum_{j in [off[a], off[a+1])} dkv[m, j] B[j, k]: one CTA per (adapter, text row)
__global__ void __launch_bounds__(128)
unet_lora_dz_kernel(const half_t* __restrict__ dkv, const float* __restrict__ B, const int* __restrict__ off,
                    float* __restrict__ dZ, int KV, int R, int r, float s) {
  __shared__ float red[4][UL_RMAX];
  const int a = blockIdx.x, m = blockIdx.y;
  float acc[UL_RMAX];
#pragma unroll
  for (int k = 0; k < UL_RMAX; ++k) acc[k] = 0.f;
  for (int j = off[a] + threadIdx.x; j < off[a + 1]; j += blockDim.x) {
    const float g = h2f(dkv[(size_t)m * KV + j]);
#pragma unroll
    for (int k = 0; k < UL_RMAX; ++k)
      if (k < r) acc[k] += g * B[(size_t)j * r + k];
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < UL_RMAX; ++k) {
    if (k < r) {
      float v = acc[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) red[warp][k] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x < r)
    dZ[(size_t)m * R + a * r + threadIdx.x] =
        s * (red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x]);
}

// dB[j, k] += s * sum_m dkv[m, j] Z[m, blk[j] r + k]: one thread per column j, the text rows split over blockIdx.y
__global__ void __launch_bounds__(128)
unet_lora_grad_b_kernel(const half_t* __restrict__ dkv, const float* __restrict__ Z, const int* __restrict__ blk,
                        float* __restrict__ dB, int M, int KV, int R, int r, float s, int rows_per_cta) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= KV) return;
  const int m0 = blockIdx.y * rows_per_cta, m1 = m0 + rows_per_cta < M ? m0 + rows_per_cta : M;
  const int zo = blk[j] * r;
  float acc[UL_RMAX];
#pragma unroll
  for (int k = 0; k < UL_RMAX; ++k) acc[k] = 0.f;
  for (int m = m0; m < m1; ++m) {
    const float g = h2f(dkv[(size_t)m * KV + j]);
    const float* z = Z + (size_t)m * R + zo;
#pragma unroll
    for (int k = 0; k < UL_RMAX; ++k)
      if (k < r) acc[k] += g * z[k];
  }
#pragma unroll
  for (int k = 0; k < UL_RMAX; ++k)
    if (k < r) atomicAdd(dB + (size_t)j * r + k, s * acc[k]);
}

// dA[q, c] += sum_m dZ[m, q] ehs[m, c]  and  d_ehs[m, c] += sum_q dZ[m, q] A[q, c]: one thread per context column c
__global__ void __launch_bounds__(256)
unet_lora_grad_a_kernel(const float* __restrict__ dZ, const half_t* __restrict__ ehs, float* __restrict__ dA, int M,
                        int ctx, int R) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, q = blockIdx.y;
  if (c >= ctx) return;
  float acc = 0.f;
  for (int m = 0; m < M; ++m) acc += dZ[(size_t)m * R + q] * h2f(ehs[(size_t)m * ctx + c]);
  dA[(size_t)q * ctx + c] += acc;
}
__global__ void __launch_bounds__(256)
unet_lora_dehs_kernel(const float* __restrict__ dZ, const float* __restrict__ A, float* __restrict__ d_ehs, int ctx,
                      int R) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
  if (c >= ctx) return;
  float acc = 0.f;
  for (int q = 0; q < R; ++q) acc += dZ[(size_t)m * R + q] * A[(size_t)q * ctx + c];
  d_ehs[(size_t)m * ctx + c] += acc;
}

}  // namespace tb

using namespace tb;

static int unet_lora_args_ok(int M, int ctx, int KV, int n_adapters, int r) {
  return M > 0 && ctx > 0 && ctx <= 4096 && KV > 0 && KV % 2 == 0 && n_adapters > 0 && r >= 1 && r <= UL_RMAX;
}

extern "C" int tb_unet_lora_fwd(const void* ehs, const float* A, const float* B, const int32_t* blk, float* Z,
                                void* kv, int M, int ctx, int KV, int n_adapters, int r, float scaling,
                                void* stream) {
  int rc = tb_check_device();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  TB_REQUIRE(ehs && A && B && blk && Z && kv && unet_lora_args_ok(M, ctx, KV, n_adapters, r), TB_E_ARG,
             "tb_unet_lora_fwd: bad args (M=%d ctx=%d KV=%d adapters=%d r=%d)", M, ctx, KV, n_adapters, r);
  const int R = n_adapters * r;
  unet_lora_down_kernel<<<M, 256, ctx * sizeof(float), st>>>((const half_t*)ehs, A, Z, ctx, R);
  if ((rc = check_launch("unet_lora_down_kernel"))) return rc;
  unet_lora_up_kernel<<<dim3((KV / 2 + 255) / 256, M), 256, 0, st>>>((half_t*)kv, Z, B, blk, KV, R, r, scaling);
  return check_launch("unet_lora_up_kernel");
}

extern "C" int tb_unet_lora_bwd(const void* dkv, const void* ehs, const float* A, const float* B, const float* Z,
                                const int32_t* blk, const int32_t* off, float* dZ, float* dA, float* dB,
                                float* d_ehs, int M, int ctx, int KV, int n_adapters, int r, float scaling,
                                void* stream) {
  int rc = tb_check_device();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  TB_REQUIRE(dkv && ehs && A && B && Z && blk && off && dZ && dA && dB && d_ehs &&
                 unet_lora_args_ok(M, ctx, KV, n_adapters, r),
             TB_E_ARG, "tb_unet_lora_bwd: bad args (M=%d ctx=%d KV=%d adapters=%d r=%d)", M, ctx, KV, n_adapters, r);
  const int R = n_adapters * r;
  unet_lora_dz_kernel<<<dim3(n_adapters, M), 128, 0, st>>>((const half_t*)dkv, B, off, dZ, KV, R, r, scaling);
  if ((rc = check_launch("unet_lora_dz_kernel"))) return rc;
  const int rows_per_cta = 80;
  unet_lora_grad_b_kernel<<<dim3((KV + 127) / 128, (M + rows_per_cta - 1) / rows_per_cta), 128, 0, st>>>(
      (const half_t*)dkv, Z, blk, dB, M, KV, R, r, scaling, rows_per_cta);
  if ((rc = check_launch("unet_lora_grad_b_kernel"))) return rc;
  unet_lora_grad_a_kernel<<<dim3((ctx + 255) / 256, R), 256, 0, st>>>(dZ, (const half_t*)ehs, dA, M, ctx, R);
  if ((rc = check_launch("unet_lora_grad_a_kernel"))) return rc;
  unet_lora_dehs_kernel<<<dim3((ctx + 255) / 256, M), 256, 0, st>>>(dZ, A, d_ehs, ctx, R);
  return check_launch("unet_lora_dehs_kernel");
}
