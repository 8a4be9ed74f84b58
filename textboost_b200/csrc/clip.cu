// Text-encoder-specific kernels (everything that is not a GEMM / LayerNorm):
//   token+position embedding gather, LoRA down-projection / weight packing / gradients, causal
//   attention over 77 tokens (forward and backward), activation forward/backward, the TextBoostModel
//   null-embedding override (textboost/text_encoder.py:71-86), sparse embedding-row gradients and the
//   knowledge-preservation loss (train_textboost.py:1096-1106).
// Sizes are tiny (616 tokens x 768): plain SIMT kernels, fp32 accumulation.
#include "host_util.h"
#include "sm100.cuh"

namespace tb {

__device__ __forceinline__ float warp_sum_c(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// x[m,:] = tok(ids[m]) + pos[m % L];  tok(id) = id < n_base ? base[id]*decay : added[id-n_base]
// An id outside [0, n_base + n_rows) (torch's embedding lookup raises for it) never dereferences: its row is
// filled with NaN, so the loss turns NaN and the GradScaler skips the step instead of reading foreign memory.
__global__ void clip_embed_kernel(const long long* __restrict__ ids, const float* __restrict__ base,
                                  const float* __restrict__ added, const float* __restrict__ decay,
                                  const float* __restrict__ pos, float* __restrict__ x, int M, int L, int D,
                                  int n_base, int n_rows) {
  const int m = blockIdx.x;
  const long long id = ids[m];
  if (id < 0 || id >= (long long)n_base + n_rows || (id >= n_base && !added)) {
    for (int c = threadIdx.x; c < D; c += blockDim.x) x[(long long)m * D + c] = nanf("");
    return;
  }
  const float dc = decay ? *decay : 1.f;
  const float* src = id < n_base ? base + id * (long long)D : added + (id - n_base) * (long long)D;
  const float sc = id < n_base ? dc : 1.f;
  const float* pp = pos + (long long)(m % L) * D;
  for (int c = threadIdx.x; c < D; c += blockDim.x) x[(long long)m * D + c] = src[c] * sc + pp[c];
}

// dE[id-n_base, :] += g[m, :] for id >= n_base (added rows only: train_textboost.py:1109-1117)
__global__ void clip_embed_grad_kernel(const long long* __restrict__ ids, const float* __restrict__ g,
                                       float* __restrict__ grad_rows, int M, int D, int n_base) {
  const int m = blockIdx.x;
  const long long id = ids[m];
  if (id < n_base) return;
  float* dst = grad_rows + (id - n_base) * (long long)D;
  for (int c = threadIdx.x; c < D; c += blockDim.x) atomicAdd(dst + c, g[(long long)m * D + c]);
}

// xa[m, j] = sum_c y[m,c] * A[j,c]  -> written as fp16 into the K-extension columns of the GEMM A operand
// (columns D .. D+R-1 of a row of stride ld; columns D+R .. D+RPAD-1 are zeroed).  R <= LORA_RMAX, processed 16
// down-projection rows at a time (the activation row is re-read from L1 per group).
constexpr int LORA_RMAX = 64;  // widest K extension: four fused targets x rank 16
__global__ void lora_down_kernel(tb::half_t* __restrict__ y_ext, long long ld, const float* __restrict__ A,
                                 int M, int D, int R, int RPAD) {
  const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= M) return;
  const int lane = threadIdx.x & 31;
  tb::half_t* row = y_ext + (long long)m * ld;
  for (int j0 = 0; j0 < RPAD; j0 += 16) {
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
    if (j0 < R) {
      // 8 consecutive columns per lane and trip: one 16-byte load of y, two float4 loads of each A row (L1-resident)
      for (int c = lane * 8; c < D; c += 256) {
        const uint4 q = *reinterpret_cast<const uint4*>(row + c);
        const tb::half2_t* h = reinterpret_cast<const tb::half2_t*>(&q);
        float v[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = tb::h22f2(h[i]);
          v[2 * i] = f.x;
          v[2 * i + 1] = f.y;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (j0 + j < R) {
            const float4 a0 = *reinterpret_cast<const float4*>(A + (long long)(j0 + j) * D + c);
            const float4 a1 = *reinterpret_cast<const float4*>(A + (long long)(j0 + j) * D + c + 4);
            acc[j] += v[0] * a0.x + v[1] * a0.y + v[2] * a0.z + v[3] * a0.w + v[4] * a1.x + v[5] * a1.y +
                      v[6] * a1.z + v[7] * a1.w;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = warp_sum_c(acc[j]);
    }
    if (lane < 16 && j0 + lane < RPAD) {
      float v = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j == lane && j0 + j < R) v = acc[j];
      row[D + j0 + lane] = tb::f2h(v);
    }
  }
}

// Pack the LoRA up-projections into the extension columns of a fused projection weight and its transpose.  The
// weight stacks `nblk` row blocks of D output features (q | k | v, or the single out_proj block); bit b of `tmask`
// says block b carries LoRA, and the i-th set bit owns B_i (Bm is [T][D][r] over the set bits, in order) and the
// extension columns [i*r, (i+1)*r):
//   Wext[b*D + n, D + i*r + j]   = scaling * B_i[n, j]        (forward operand, [nblk*D, D+RPAD])
//   WextT[D + i*r + j, b*D + n]  = scaling * B_i[n, j]        (dgrad operand,  [D+RPAD, nblk*D])
// everything else inside the extension block is zero.
__device__ __forceinline__ int lora_bits(int m) {  // (<= 8 blocks; plain arithmetic so the source also builds for the host)
  int c = 0;
  for (int b = 0; b < 8; ++b) c += (m >> b) & 1;
  return c;
}
__device__ __forceinline__ int lora_slot(int tmask, int b) {  // index among the set bits, or -1
  return ((tmask >> b) & 1) ? lora_bits(tmask & ((1 << b) - 1)) : -1;
}
__global__ void lora_pack_kernel(const float* __restrict__ Bm, tb::half_t* __restrict__ Wext,
                                 tb::half_t* __restrict__ WextT, int nblk, int tmask, int D, int r, int RPAD,
                                 float scaling) {
  const int K = D + RPAD;
  const long long total = (long long)nblk * D * RPAD;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i % RPAD);
    const long long row = i / RPAD;  // b*D + n
    const int b = (int)(row / D), n = (int)(row % D);
    const int t = lora_slot(tmask, b);
    float v = 0.f;
    if (t >= 0 && e >= t * r && e < (t + 1) * r) v = scaling * Bm[((long long)t * D + n) * r + (e - t * r)];
    const tb::half_t hv = tb::f2h(v);
    Wext[row * K + D + e] = hv;
    WextT[(long long)(D + e) * (nblk * D) + row] = hv;
  }
}

// LoRA gradients for one fused projection of one layer (accumulating, fp32):
//   dB_i[n, j] += scaling * sum_m dY[m, b_i*D+n] * xa[m, i*r+j]        (b_i = the i-th set bit of tmask)
//   dA[ij, c]  += sum_m dxa[m, ij] * y[m, c]
// dY [M, nblk*D] fp16, xa = y_ext[:, D:D+R], y = y_ext[:, :D], dxa = dA_ext[:, D:D+R] (all fp16), R = T*r <= 64.
// A thread owns two adjacent columns (of dY for dB, of y for dA), so every warp reads 128 contiguous bytes per
// row; the rows are split over grid.y and the partial sums meet in fp32 atomics; the R-wide xa / dxa rows of the
// CTA's row chunk are staged in shared memory (every thread reads the same entry: broadcast).
constexpr int LG_ROWS = 16;  // rows per CTA: all of a thread's row loads are in flight at once
// Thread = 2 consecutive columns (one half2 per row, a warp reads 128 contiguous bytes) x LG_ROWS rows.  Columns
// [0, nblk*D) are outputs of the fused projection (dB of the LoRA target that owns them), columns
// [nblk*D, nblk*D + D) the LayerNorm output (dA).  The reduction over the M = batch x 77 rows is split over grid.y
// and finished with fp32 atomics.  Shape of the problem: 4 MB of operands, 11 MFMA -- what matters is that enough
// warps are resident to hide the load latency (a first version with 48 serial rows per thread took 26 us, one with 8
// columns x 16 rows per thread and 206 registers 21 us at 6 % occupancy; ncu, profiles/r02_ncu_cases_summary.txt).
__global__ void __launch_bounds__(128)
lora_grad_kernel(const tb::half_t* __restrict__ dY, const tb::half_t* __restrict__ y_ext,
                 const tb::half_t* __restrict__ dA_ext, long long ld, float* __restrict__ dB,
                 float* __restrict__ dA, int M, int nblk, int tmask, int D, int r, float scaling) {
  __shared__ float sxa[LG_ROWS][LORA_RMAX];
  __shared__ float sdx[LG_ROWS][LORA_RMAX];
  const int m0 = blockIdx.y * LG_ROWS;
  const int rows = min(LG_ROWS, M - m0);
  const int R = lora_bits(tmask) * r;
  const int NY = nblk * D;
  const int col = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
  const bool is_b = col < NY, live = col < NY + D;
  const int t = is_b ? lora_slot(tmask, col / D) : 0;
  const bool work = live && t >= 0;
  // this thread's operand rows first: their latency overlaps the staging of the coefficient rows below
  tb::half2_t q[LG_ROWS];
  if (work) {
    const tb::half_t* src = is_b ? dY + (long long)m0 * NY + col : y_ext + (long long)m0 * ld + (col - NY);
    const long long lds = is_b ? NY : ld;
#pragma unroll
    for (int mm = 0; mm < LG_ROWS; ++mm)
      q[mm] = mm < rows ? *reinterpret_cast<const tb::half2_t*>(src + (long long)mm * lds) : tb::half2_t{};
  }
  // coefficient rows: the R extension columns of y_ext (xa) and dA_ext (dxa), two per load
  for (int i = threadIdx.x; i < LG_ROWS * (LORA_RMAX / 2); i += blockDim.x) {
    const int mm = i / (LORA_RMAX / 2), j = 2 * (i % (LORA_RMAX / 2));
    float2 a = {0.f, 0.f}, b = {0.f, 0.f};
    if (mm < rows && j < R) {  // (R is even or the K extension is padded with zeros: reading j + 1 is in bounds)
      a = tb::h22f2(*reinterpret_cast<const tb::half2_t*>(y_ext + (long long)(m0 + mm) * ld + D + j));
      b = tb::h22f2(*reinterpret_cast<const tb::half2_t*>(dA_ext + (long long)(m0 + mm) * ld + D + j));
      if (j + 1 >= R) a.y = b.y = 0.f;
    }
    sxa[mm][j] = a.x;
    sxa[mm][j + 1] = a.y;
    sdx[mm][j] = b.x;
    sdx[mm][j + 1] = b.y;
  }
  __syncthreads();
  if (!work) return;
  const int jn = is_b ? r : R;            // coefficients this thread contracts with
  const int jbase = is_b ? t * r : 0;
  float v0[LG_ROWS], v1[LG_ROWS];
#pragma unroll
  for (int mm = 0; mm < LG_ROWS; ++mm) {
    const float2 f = tb::h22f2(q[mm]);
    v0[mm] = f.x;
    v1[mm] = f.y;
  }
  for (int j = 0; j < jn; ++j) {
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int mm = 0; mm < LG_ROWS; ++mm) {
      const float w = is_b ? sxa[mm][jbase + j] : sdx[mm][jbase + j];
      a0 += v0[mm] * w;
      a1 += v1[mm] * w;
    }
    if (is_b) {
      float* dst = dB + ((long long)t * D + col % D) * r;  // D even: both columns inside one target block
      atomicAdd(&dst[j], scaling * a0);
      atomicAdd(&dst[r + j], scaling * a1);
    } else {
      float* dst = dA + (long long)j * D + (col - NY);
      atomicAdd(&dst[0], a0);
      atomicAdd(&dst[1], a1);
    }
  }
}

// dy[m, c] += sum_j dxa[m, j] * A[j, c]   (in place on the first D columns of dA_ext)
__global__ void lora_dx_kernel(tb::half_t* __restrict__ dA_ext, long long ld, const float* __restrict__ A,
                               int M, int D, int R) {
  const int m = blockIdx.x;
  __shared__ float sx[LORA_RMAX];
  if (threadIdx.x < LORA_RMAX) sx[threadIdx.x] = threadIdx.x < R ? tb::h2f(dA_ext[(long long)m * ld + D + threadIdx.x]) : 0.f;
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float v = tb::h2f(dA_ext[(long long)m * ld + c]);
    for (int j = 0; j < R; ++j) v += sx[j] * A[(long long)j * D + c];
    dA_ext[(long long)m * ld + c] = tb::f2h(v);
  }
}

// ------------------------------------------------------------------ causal attention, L <= 128, d = 64
// qkv [B*L, 3*D] fp16 (q | k | v, head h at columns h*64); one CTA per (b, h).
template <bool BWD>
__global__ void clip_attn_kernel(const tb::half_t* __restrict__ qkv, const tb::half_t* __restrict__ dO,
                                 tb::half_t* __restrict__ out, int L, int D, int heads, float scale) {
  constexpr int HD = 64;
  extern __shared__ float smf[];
  const int PS = (L * (L + 1) + 3) & ~3;  // keep the fp16 tiles 16-byte aligned
  float* sP = smf;                        // [L][L+1]
  float* sdS = sP + PS;                   // [L][L+1] (BWD only)
  tb::half_t* sQ = reinterpret_cast<tb::half_t*>(BWD ? sdS + PS : sP + PS);
  tb::half_t* sK = sQ + L * HD;
  tb::half_t* sV = sK + L * HD;
  tb::half_t* sdO = sV + L * HD;  // BWD only
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const long long rs = 3LL * D;
  const tb::half_t* base = qkv + (long long)b * L * rs + h * HD;
  for (int i = threadIdx.x; i < L * HD / 8; i += blockDim.x) {
    const int l = i / (HD / 8), v = i % (HD / 8);
    reinterpret_cast<uint4*>(sQ)[i] = *reinterpret_cast<const uint4*>(base + l * rs + v * 8);
    reinterpret_cast<uint4*>(sK)[i] = *reinterpret_cast<const uint4*>(base + l * rs + D + v * 8);
    reinterpret_cast<uint4*>(sV)[i] = *reinterpret_cast<const uint4*>(base + l * rs + 2 * D + v * 8);
    if (BWD)
      reinterpret_cast<uint4*>(sdO)[i] =
          *reinterpret_cast<const uint4*>(dO + ((long long)b * L + l) * D + h * HD + v * 8);
  }
  __syncthreads();
  // S = scale * Q K^T with the causal mask (j <= i)
  for (int idx = threadIdx.x; idx < L * L; idx += blockDim.x) {
    const int i = idx / L, j = idx % L;
    float acc = -INFINITY;
    if (j <= i) {
      acc = 0.f;
      const tb::half2_t* qp = reinterpret_cast<const tb::half2_t*>(sQ + i * HD);
      const tb::half2_t* kp = reinterpret_cast<const tb::half2_t*>(sK + j * HD);
#pragma unroll 8
      for (int c = 0; c < HD / 2; ++c) {
        const float2 a = tb::h22f2(qp[c]), bb = tb::h22f2(kp[c]);
        acc += a.x * bb.x + a.y * bb.y;
      }
      acc *= scale;
    }
    sP[i * (L + 1) + j] = acc;
  }
  __syncthreads();
  // row softmax (fp32), one warp per row
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int i = warp; i < L; i += nw) {
    float mx = -INFINITY;
    for (int j = lane; j <= i; j += 32) mx = fmaxf(mx, sP[i * (L + 1) + j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int j = lane; j < L; j += 32) {
      const float e = j <= i ? __expf(sP[i * (L + 1) + j] - mx) : 0.f;
      sP[i * (L + 1) + j] = e;
      sum += e;
    }
    sum = warp_sum_c(sum);
    const float inv = 1.f / sum;
    // the reference casts the probabilities to the matmul dtype (fp16 under autocast) before P V
    for (int j = lane; j < L; j += 32) sP[i * (L + 1) + j] = tb::h2f(tb::f2h(sP[i * (L + 1) + j] * inv));
  }
  __syncthreads();
  if (!BWD) {
    for (int idx = threadIdx.x; idx < L * HD; idx += blockDim.x) {
      const int i = idx / HD, c = idx % HD;
      float acc = 0.f;
      for (int j = 0; j <= i; ++j) acc += sP[i * (L + 1) + j] * tb::h2f(sV[j * HD + c]);
      out[((long long)b * L + i) * D + h * HD + c] = tb::f2h(acc);
    }
    return;
  }
  // backward: dP = dO V^T ; dS = P o (dP - rowsum(P o dP)) ; dQ = scale dS K ; dK = scale dS^T Q ; dV = P^T dO
  for (int idx = threadIdx.x; idx < L * L; idx += blockDim.x) {
    const int i = idx / L, j = idx % L;
    float acc = 0.f;
    if (j <= i) {
      const tb::half2_t* gp = reinterpret_cast<const tb::half2_t*>(sdO + i * HD);
      const tb::half2_t* vp = reinterpret_cast<const tb::half2_t*>(sV + j * HD);
#pragma unroll 8
      for (int c = 0; c < HD / 2; ++c) {
        const float2 a = tb::h22f2(gp[c]), bb = tb::h22f2(vp[c]);
        acc += a.x * bb.x + a.y * bb.y;
      }
    }
    sdS[i * (L + 1) + j] = acc;
  }
  __syncthreads();
  for (int i = warp; i < L; i += nw) {
    float dsum = 0.f;
    for (int j = lane; j <= i; j += 32) dsum += sP[i * (L + 1) + j] * sdS[i * (L + 1) + j];
    dsum = warp_sum_c(dsum);
    for (int j = lane; j < L; j += 32)
      sdS[i * (L + 1) + j] = j <= i ? sP[i * (L + 1) + j] * (sdS[i * (L + 1) + j] - dsum) * scale : 0.f;
  }
  __syncthreads();
  tb::half_t* obase = out + (long long)b * L * rs + h * HD;  // d(qkv) laid out like qkv
  for (int idx = threadIdx.x; idx < L * HD; idx += blockDim.x) {
    const int i = idx / HD, c = idx % HD;
    float dq = 0.f, dk = 0.f, dv = 0.f;
    for (int j = 0; j <= i; ++j) dq += sdS[i * (L + 1) + j] * tb::h2f(sK[j * HD + c]);
    for (int j = i; j < L; ++j) {
      dk += sdS[j * (L + 1) + i] * tb::h2f(sQ[j * HD + c]);
      dv += sP[j * (L + 1) + i] * tb::h2f(sdO[j * HD + c]);
    }
    obase[i * rs + c] = tb::f2h(dq);
    obase[i * rs + D + c] = tb::f2h(dk);
    obase[i * rs + 2 * D + c] = tb::f2h(dv);
  }
}

// ------------------------------------------------------------------ activations (fc1 -> act -> fc2)
__device__ __forceinline__ float act_fwd(float u, int kind) {
  if (kind == TB_ACT_QUICK_GELU) return __fdividef(u, 1.f + __expf(-1.702f * u));
  return 0.5f * u * (1.f + erff(u * 0.70710678118654752f));
}
__device__ __forceinline__ float act_bwd(float u, int kind) {
  if (kind == TB_ACT_QUICK_GELU) {
    const float s = __fdividef(1.f, 1.f + __expf(-1.702f * u));
    return s * (1.f + 1.702f * u * (1.f - s));
  }
  const float cdf = 0.5f * (1.f + erff(u * 0.70710678118654752f));
  return cdf + u * 0.3989422804014327f * __expf(-0.5f * u * u);
}
// BWD=false: out = act(u);  BWD=true: out = g * act'(u).  Eight elements per thread and trip (16-byte accesses).
template <bool BWD>
__global__ void act_kernel(const tb::half_t* __restrict__ u, const tb::half_t* __restrict__ g,
                           tb::half_t* __restrict__ out, long long n, int kind) {
  const long long nv = n / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv;
       i += (long long)gridDim.x * blockDim.x) {
    const uint4 qu = *reinterpret_cast<const uint4*>(u + i * 8);
    uint4 qg = make_uint4(0u, 0u, 0u, 0u);
    if (BWD) qg = *reinterpret_cast<const uint4*>(g + i * 8);
    const tb::half_t* hu = reinterpret_cast<const tb::half_t*>(&qu);
    const tb::half_t* hg = reinterpret_cast<const tb::half_t*>(&qg);
    uint4 qo;
    tb::half_t* ho = reinterpret_cast<tb::half_t*>(&qo);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float x = tb::h2f(hu[k]);
      ho[k] = tb::f2h(BWD ? tb::h2f(hg[k]) * act_bwd(x, kind) : act_fwd(x, kind));
    }
    *reinterpret_cast<uint4*>(out + i * 8) = qo;
  }
  // tail (n % 8 elements), one thread
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = nv * 8; i < n; ++i) {
      const float x = tb::h2f(u[i]);
      out[i] = tb::f2h(BWD ? tb::h2f(g[i]) * act_bwd(x, kind) : act_fwd(x, kind));
    }
}

// ------------------------------------------------------------------ TextBoostModel override
// forward : rows with ids[b,1]==eos -> null[l,:]; if fixed: position 0 -> null[0,:]   (text_encoder.py:71-86)
// backward: the overwritten slots receive no gradient -> zero them.
template <bool BWD>
__global__ void null_override_kernel(const long long* __restrict__ ids, const float* __restrict__ null_emb,
                                     float* __restrict__ h, int L, int D, int eos_id, int use_fixed) {
  const int m = blockIdx.x;
  const int b = m / L, l = m % L;
  const bool whole = ids[(long long)b * L + 1] == eos_id;
  const bool hit = whole || (use_fixed && l == 0);
  if (!hit) return;
  for (int c = threadIdx.x; c < D; c += blockDim.x)
    h[(long long)m * D + c] = BWD ? 0.f : null_emb[(long long)l * D + c];
}

// fp32 [M,D] -> fp16 copy (encoder_hidden_states.to(unet.dtype), train_textboost.py:1066)
// ------------------------------------------------------------------ knowledge-preservation loss
// cos : loss += w * mean_m (1 - cos(h_m, h0_m));  dh = -w/M * ls * (h0/(|h||h0|) - cos h/|h|^2)
// mse : loss += w * mean((h-h0)^2);               dh = w * 2 (h-h0)/(M*D) * ls
// dh is ACCUMULATED into dh_acc (the same buffer also receives d ehs from the UNet for other rows).
__global__ void kpl_kernel(const float* __restrict__ h, const float* __restrict__ h0, int M, int D, int kind,
                           float weight, const float* __restrict__ loss_scale, float* __restrict__ loss_acc,
                           float* __restrict__ dh) {
  const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= M) return;
  const int lane = threadIdx.x & 31;
  const float ls = loss_scale ? *loss_scale : 1.f;
  const float* a = h + (long long)m * D;
  const float* b = h0 + (long long)m * D;
  float* g = dh ? dh + (long long)m * D : nullptr;
  if (kind == 0) {
    float ab = 0.f, aa = 0.f, bb = 0.f;
    for (int c = lane; c < D; c += 32) {
      ab += a[c] * b[c];
      aa += a[c] * a[c];
      bb += b[c] * b[c];
    }
    ab = warp_sum_c(ab);
    aa = warp_sum_c(aa);
    bb = warp_sum_c(bb);
    // F.cosine_similarity clamps each norm at eps = 1e-8
    const float na = fmaxf(sqrtf(aa), 1e-8f), nb = fmaxf(sqrtf(bb), 1e-8f);
    const float cs = ab / (na * nb);
    if (lane == 0) atomicAdd(loss_acc, weight * (1.f - cs) / M);
    if (g) {
      const float k = -weight / M * ls;
      for (int c = lane; c < D; c += 32) g[c] += k * (b[c] / (na * nb) - cs * a[c] / (na * na));
    }
  } else {
    float s = 0.f;
    const float k = weight * 2.f / ((float)M * D) * ls;
    for (int c = lane; c < D; c += 32) {
      const float d = a[c] - b[c];
      s += d * d;
      if (g) g[c] += k * d;
    }
    s = warp_sum_c(s);
    if (lane == 0) atomicAdd(loss_acc, weight * s / ((float)M * D));
  }
}

}  // namespace tb

using namespace tb;

#define TB_ENTER()            \
  int rc = tb_check_device(); \
  if (rc) return rc;          \
  cudaStream_t st = (cudaStream_t)stream

extern "C" int tb_clip_embed(const int64_t* ids, const float* base, const float* added, const float* decay,
                             const float* pos, float* x, int M, int L, int D, int n_base, int n_rows,
                             void* stream) {
  TB_ENTER();
  TB_REQUIRE(ids && base && pos && x, TB_E_ARG, "tb_clip_embed: null pointer");
  TB_REQUIRE(n_rows >= 0 && (n_rows == 0 || added), TB_E_ARG, "tb_clip_embed: n_rows = %d without added rows", n_rows);
  clip_embed_kernel<<<M, 256, 0, st>>>((const long long*)ids, base, added, decay, pos, x, M, L, D, n_base, n_rows);
  return check_launch("clip_embed_kernel");
}
extern "C" int tb_clip_embed_grad(const int64_t* ids, const float* g, float* grad_rows, int M, int D,
                                  int n_base, void* stream) {
  TB_ENTER();
  TB_REQUIRE(ids && g && grad_rows, TB_E_ARG, "tb_clip_embed_grad: null pointer");
  clip_embed_grad_kernel<<<M, 256, 0, st>>>((const long long*)ids, g, grad_rows, M, D, n_base);
  return check_launch("clip_embed_grad_kernel");
}
extern "C" int tb_lora_down(void* y_ext, int64_t ld, const float* A, int M, int D, int R, int RPAD,
                            void* stream) {
  TB_ENTER();
  TB_REQUIRE(y_ext && A && R >= 1 && R <= LORA_RMAX && RPAD <= LORA_RMAX && R <= RPAD && RPAD % 8 == 0, TB_E_ARG,
             "tb_lora_down: bad args (R=%d RPAD=%d; R <= RPAD <= %d)", R, RPAD, LORA_RMAX);
  TB_REQUIRE(D % 8 == 0 && ld % 8 == 0, TB_E_ALIGN, "tb_lora_down: D and ld must be multiples of 8");
  lora_down_kernel<<<(M + 7) / 8, 256, 0, st>>>((tb::half_t*)y_ext, ld, A, M, D, R, RPAD);
  return check_launch("lora_down_kernel");
}
static int lora_mask_ok(int nblk, int tmask, int r, int RPAD) {
  if (nblk < 1 || nblk > 8 || tmask <= 0 || tmask >= (1 << nblk) || r < 1 || r > 16) return 0;
  int T = 0;
  for (int b = 0; b < nblk; ++b) T += (tmask >> b) & 1;
  return T * r <= RPAD && RPAD <= LORA_RMAX;
}
extern "C" int tb_lora_pack(const float* Bm, void* Wext, void* WextT, int nblk, int tmask, int D, int r, int RPAD,
                            float scaling, void* stream) {
  TB_ENTER();
  TB_REQUIRE(Bm && Wext && WextT && lora_mask_ok(nblk, tmask, r, RPAD), TB_E_ARG,
             "tb_lora_pack: bad args (nblk=%d tmask=%d r=%d RPAD=%d)", nblk, tmask, r, RPAD);
  const long long total = (long long)nblk * D * RPAD;
  lora_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(Bm, (tb::half_t*)Wext, (tb::half_t*)WextT, nblk, tmask,
                                                                   D, r, RPAD, scaling);
  return check_launch("lora_pack_kernel");
}
extern "C" int tb_lora_grad(const void* dY, const void* y_ext, const void* dA_ext, int64_t ld, float* dB,
                            float* dA, int M, int nblk, int tmask, int D, int r, float scaling, void* stream) {
  TB_ENTER();
  TB_REQUIRE(dY && y_ext && dA_ext && dB && dA && lora_mask_ok(nblk, tmask, r, LORA_RMAX), TB_E_ARG,
             "tb_lora_grad: bad args (nblk=%d tmask=%d r=%d)", nblk, tmask, r);
  TB_REQUIRE(D % 2 == 0 && ld % 2 == 0, TB_E_ALIGN, "tb_lora_grad: D and ld must be even");
  const int pairs = (nblk * D + D) / 2;
  dim3 grid((pairs + 127) / 128, (M + LG_ROWS - 1) / LG_ROWS);
  lora_grad_kernel<<<grid, 128, 0, st>>>((const tb::half_t*)dY, (const tb::half_t*)y_ext, (const tb::half_t*)dA_ext, ld, dB, dA,
                                         M, nblk, tmask, D, r, scaling);
  return check_launch("lora_grad_kernel");
}
extern "C" int tb_lora_dx(void* dA_ext, int64_t ld, const float* A, int M, int D, int R, void* stream) {
  TB_ENTER();
  TB_REQUIRE(dA_ext && A && R >= 1 && R <= LORA_RMAX, TB_E_ARG, "tb_lora_dx: bad args (R=%d)", R);
  lora_dx_kernel<<<M, 256, 0, st>>>((tb::half_t*)dA_ext, ld, A, M, D, R);
  return check_launch("lora_dx_kernel");
}
extern "C" int tb_clip_attn_fwd(const void* qkv, void* out, int B, int L, int D, int heads, void* stream) {
  TB_ENTER();
  TB_REQUIRE(qkv && out && D == heads * 64 && L <= 128, TB_E_SHAPE,
             "tb_clip_attn_fwd: head_dim must be 64 and L <= 128 (D=%d heads=%d L=%d)", D, heads, L);
  const int smem = ((L * (L + 1) + 3) & ~3) * 4 + 3 * L * 64 * 2;
  static bool cfg = false;
  if (!cfg) {
    cudaFuncSetAttribute(clip_attn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 129 * 4 + 3 * 128 * 64 * 2);
    cfg = true;
  }
  clip_attn_kernel<false><<<B * heads, 256, smem, st>>>((const tb::half_t*)qkv, nullptr, (tb::half_t*)out, L, D,
                                                        heads, 0.125f);
  return check_launch("clip_attn_kernel<fwd>");
}
extern "C" int tb_clip_attn_bwd(const void* qkv, const void* dO, void* dqkv, int B, int L, int D, int heads,
                                void* stream) {
  TB_ENTER();
  TB_REQUIRE(qkv && dO && dqkv && D == heads * 64 && L <= 128, TB_E_SHAPE, "tb_clip_attn_bwd: bad shape");
  const int smem = 2 * ((L * (L + 1) + 3) & ~3) * 4 + 4 * L * 64 * 2;
  static bool cfg = false;
  if (!cfg) {
    cudaFuncSetAttribute(clip_attn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 128 * 129 * 4 + 4 * 128 * 64 * 2);
    cfg = true;
  }
  clip_attn_kernel<true><<<B * heads, 256, smem, st>>>((const tb::half_t*)qkv, (const tb::half_t*)dO, (tb::half_t*)dqkv,
                                                       L, D, heads, 0.125f);
  return check_launch("clip_attn_kernel<bwd>");
}
extern "C" int tb_act_fwd_f16(const void* u, void* out, int64_t n, int kind, void* stream) {
  TB_ENTER();
  TB_REQUIRE(u && out && (kind == TB_ACT_QUICK_GELU || kind == TB_ACT_GELU), TB_E_ARG, "tb_act_fwd_f16: bad args");
  TB_REQUIRE(((uintptr_t)u | (uintptr_t)out) % 16 == 0, TB_E_ALIGN, "tb_act_fwd_f16: 16-byte alignment");
  act_kernel<false><<<(unsigned)((n / 8 + 255) / 256 > 1184 ? 1184 : (n / 8 + 255) / 256 + 1), 256, 0, st>>>(
      (const tb::half_t*)u, nullptr, (tb::half_t*)out, n, kind);
  return check_launch("act_kernel<fwd>");
}
extern "C" int tb_act_bwd_f16(const void* u, const void* g, void* out, int64_t n, int kind, void* stream) {
  TB_ENTER();
  TB_REQUIRE(u && g && out && (kind == TB_ACT_QUICK_GELU || kind == TB_ACT_GELU), TB_E_ARG, "tb_act_bwd_f16: bad args");
  TB_REQUIRE(((uintptr_t)u | (uintptr_t)g | (uintptr_t)out) % 16 == 0, TB_E_ALIGN, "tb_act_bwd_f16: 16-byte alignment");
  act_kernel<true><<<(unsigned)((n / 8 + 255) / 256 > 1184 ? 1184 : (n / 8 + 255) / 256 + 1), 256, 0, st>>>(
      (const tb::half_t*)u, (const tb::half_t*)g, (tb::half_t*)out, n, kind);
  return check_launch("act_kernel<bwd>");
}
extern "C" int tb_null_override(const int64_t* ids, const float* null_emb, float* h, int B, int L, int D,
                                int eos_id, int use_fixed, int backward, void* stream) {
  TB_ENTER();
  TB_REQUIRE(ids && h && (backward || null_emb), TB_E_ARG, "tb_null_override: null pointer");
  if (backward)
    null_override_kernel<true><<<B * L, 256, 0, st>>>((const long long*)ids, null_emb, h, L, D, eos_id, use_fixed);
  else
    null_override_kernel<false><<<B * L, 256, 0, st>>>((const long long*)ids, null_emb, h, L, D, eos_id, use_fixed);
  return check_launch("null_override_kernel");
}
extern "C" int tb_kpl_fwd_bwd(const float* h, const float* h0, int M, int D, int kind, float weight,
                              const float* loss_scale, float* loss_acc, float* dh_acc, void* stream) {
  TB_ENTER();
  TB_REQUIRE(h && h0 && loss_acc && (kind == 0 || kind == 1), TB_E_ARG, "tb_kpl_fwd_bwd: bad args");
  kpl_kernel<<<(M + 7) / 8, 256, 0, st>>>(h, h0, M, D, kind, weight, loss_scale, loss_acc, dh_acc);
  return check_launch("kpl_kernel");
}
