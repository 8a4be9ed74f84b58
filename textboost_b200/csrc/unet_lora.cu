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

// Z[m, q] = sum_c ehs[m, c] A[q, c]: one CTA per UL_DOWN_ROWS text rows (cached in shared memory as fp32), one warp per
// q -- each A row is read once per CTA and used for all its text rows
constexpr int UL_DOWN_ROWS = 4;
__global__ void __launch_bounds__(256)
unet_lora_down_kernel(const half_t* __restrict__ ehs, const float* __restrict__ A, float* __restrict__ Z, int M, int ctx,
                      int R) {
  extern __shared__ float smf[];  // [UL_DOWN_ROWS][ctx]
  const int m0 = blockIdx.x * UL_DOWN_ROWS;
  for (int i = threadIdx.x; i < UL_DOWN_ROWS * ctx; i += blockDim.x) {
    const int m = m0 + i / ctx;
    smf[i] = m < M ? h2f(ehs[(size_t)m * ctx + i % ctx]) : 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  // the adapter rows q are split over gridDim.y CTAs (154 CTAs of 8 warps would leave the machine a third full)
  for (int q = blockIdx.y * nwarp + warp; q < R; q += nwarp * gridDim.y) {
    const float* a = A + (size_t)q * ctx;
    float acc[UL_DOWN_ROWS];
#pragma unroll
    for (int i = 0; i < UL_DOWN_ROWS; ++i) acc[i] = 0.f;
    for (int c = lane; c < ctx; c += 32) {
      const float av = a[c];
#pragma unroll
      for (int i = 0; i < UL_DOWN_ROWS; ++i) acc[i] += smf[i * ctx + c] * av;
    }
#pragma unroll
    for (int i = 0; i < UL_DOWN_ROWS; ++i) {
      float v = acc[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && m0 + i < M) Z[(size_t)(m0 + i) * R + q] = v;
    }
  }
}

// The three kernels below stream the [M, KV] 16-bit tensor once, eight adjacent columns (one 16-byte vector) per
// thread; adapter column ranges are multiples of 8 (channel counts are), so a vector never straddles two adapters.
__device__ __forceinline__ void ul_unpack8(const uint4& q, float* f) {
  const half2_t* h = reinterpret_cast<const half2_t*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 v = h22f2(h[i]);
    f[2 * i] = v.x;
    f[2 * i + 1] = v.y;
  }
}

// Both kernels keep the 8 x r lora_B floats of a thread's column vector in registers (float4 loads: 8r contiguous
// floats, 32r-byte aligned) and reuse them over ROWS text rows: lora_B is read M / ROWS times instead of M times.
// RR = compile-time bound of the rank (register arrays), ROWS x RR <= 32.
template <int RR>
__device__ __forceinline__ void ul_load_b(const float* __restrict__ B, int j, int r, float (&bb)[8][RR]) {
  const float4* src = reinterpret_cast<const float4*>(B + (size_t)j * r);
  float flat[8 * RR];
#pragma unroll
  for (int q = 0; q < 2 * RR; ++q) {
    if (q < 2 * r) {
      const float4 t = src[q];
      flat[4 * q] = t.x; flat[4 * q + 1] = t.y; flat[4 * q + 2] = t.z; flat[4 * q + 3] = t.w;
    }
  }
  if (r == RR) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int k = 0; k < RR; ++k) bb[i][k] = flat[i * RR + k];
  } else {  // ranks between the template bounds: re-read the few floats with their true stride
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int k = 0; k < RR; ++k) bb[i][k] = k < r ? B[(size_t)(j + i) * r + k] : 0.f;
  }
}

// kv[m, j] += s * Z[m, blk[j] r ..] . B[j, :]
template <int RR, int ROWS>
__global__ void __launch_bounds__(128)
unet_lora_up_kernel(half_t* __restrict__ kv, const float* __restrict__ Z, const float* __restrict__ B,
                    const int* __restrict__ blk, int M, int KV, int R, int r, float s) {
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (j >= KV) return;
  float bb[8][RR];
  ul_load_b<RR>(B, j, r, bb);
  const int zo = blk[j] * r;
  const int m0 = blockIdx.y * ROWS;
#pragma unroll
  for (int mi = 0; mi < ROWS; ++mi) {
    const int m = m0 + mi;
    if (m >= M) break;
    uint4* p = reinterpret_cast<uint4*>(kv + (size_t)m * KV + j);
    float v[8], z[RR];
    ul_unpack8(*p, v);
#pragma unroll
    for (int k = 0; k < RR; ++k) z[k] = k < r ? Z[(size_t)m * R + zo + k] : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < RR; ++k) a += z[k] * bb[i][k];
      v[i] += s * a;
    }
    uint4 o;
    half2_t* oh = reinterpret_cast<half2_t*>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) oh[i] = ff2h2(v[2 * i], v[2 * i + 1]);
    *p = o;
  }
}

// dZ[m, a r + k] = s * sum_{j in [off[a], off[a+1])} dkv[m, j] B[j, k]: one CTA per (adapter, ROWS text rows)
template <int RR, int ROWS>
__global__ void __launch_bounds__(128)
unet_lora_dz_kernel(const half_t* __restrict__ dkv, const float* __restrict__ B, const int* __restrict__ off,
                    float* __restrict__ dZ, int M, int KV, int R, int r, float s) {
  __shared__ float red[4][ROWS][RR];
  const int a = blockIdx.x, m0 = blockIdx.y * ROWS;
  float acc[ROWS][RR];
#pragma unroll
  for (int mi = 0; mi < ROWS; ++mi)
#pragma unroll
    for (int k = 0; k < RR; ++k) acc[mi][k] = 0.f;
  for (int j = off[a] + threadIdx.x * 8; j < off[a + 1]; j += blockDim.x * 8) {
    float bb[8][RR];
    ul_load_b<RR>(B, j, r, bb);
#pragma unroll
    for (int mi = 0; mi < ROWS; ++mi) {
      if (m0 + mi < M) {
        float g[8];
        ul_unpack8(*reinterpret_cast<const uint4*>(dkv + (size_t)(m0 + mi) * KV + j), g);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int k = 0; k < RR; ++k) acc[mi][k] += g[i] * bb[i][k];
      }
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int mi = 0; mi < ROWS; ++mi)
#pragma unroll
    for (int k = 0; k < RR; ++k) {
      float v = acc[mi][k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) red[warp][mi][k] = v;
    }
  __syncthreads();
  for (int t = threadIdx.x; t < ROWS * r; t += blockDim.x) {
    const int mi = t / r, k = t % r;
    if (m0 + mi < M)
      dZ[(size_t)(m0 + mi) * R + a * r + k] = s * (red[0][mi][k] + red[1][mi][k] + red[2][mi][k] + red[3][mi][k]);
  }
}

// dB[j, k] += s * sum_m dkv[m, j] Z[m, blk[j] r + k]: one thread per 8 columns, the text rows split over blockIdx.y.
// RR = compile-time rank bound (accumulators stay in registers: 8 x RR floats)
template <int RR>
__global__ void __launch_bounds__(128)
unet_lora_grad_b_kernel(const half_t* __restrict__ dkv, const float* __restrict__ Z, const int* __restrict__ blk,
                        float* __restrict__ dB, int M, int KV, int R, int r, float s, int rows_per_cta) {
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (j >= KV) return;
  const int m0 = blockIdx.y * rows_per_cta, m1 = m0 + rows_per_cta < M ? m0 + rows_per_cta : M;
  const int zo = blk[j] * r;
  float acc[8][RR];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int k = 0; k < RR; ++k) acc[i][k] = 0.f;
  for (int m = m0; m < m1; ++m) {
    float g[8];
    ul_unpack8(*reinterpret_cast<const uint4*>(dkv + (size_t)m * KV + j), g);
    const float* z = Z + (size_t)m * R + zo;
#pragma unroll
    for (int k = 0; k < RR; ++k) {
      if (k < r) {
        const float zk = z[k];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i][k] += g[i] * zk;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int k = 0; k < RR; ++k)
      if (k < r) atomicAdd(dB + (size_t)(j + i) * r + k, s * acc[i][k]);
}

// dA[q, c] += sum_m dZ[m, q] ehs[m, c]  and  d_ehs[m, c] += sum_q dZ[m, q] A[q, c]: one thread per context column c.
// grad_a: eight adapter rows q per thread (the ehs element is loaded once for the eight), text rows split into
// n_chunks (folded into blockIdx.y) and combined with atomics (98 k outputs x 616 rows would otherwise be 384 CTAs of serial loops)
__global__ void __launch_bounds__(256)
unet_lora_grad_a_kernel(const float* __restrict__ dZ, const half_t* __restrict__ ehs, float* __restrict__ dA, int M,
                        int ctx, int R, int rows_per_cta, int n_chunks) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, q0 = (blockIdx.y / n_chunks) * 8;
  if (c >= ctx) return;
  const int m0 = (blockIdx.y % n_chunks) * rows_per_cta, m1 = m0 + rows_per_cta < M ? m0 + rows_per_cta : M;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int m = m0; m < m1; ++m) {
    const float e = h2f(ehs[(size_t)m * ctx + c]);
    const float* dz = dZ + (size_t)m * R + q0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (q0 + i < R) acc[i] += dz[i] * e;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (q0 + i < R) atomicAdd(dA + (size_t)(q0 + i) * ctx + c, acc[i]);
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
  return M > 0 && ctx > 0 && ctx <= 2048 && KV > 0 && KV % 8 == 0 && n_adapters > 0 && r >= 1 && r <= UL_RMAX;
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
  unet_lora_down_kernel<<<dim3((M + UL_DOWN_ROWS - 1) / UL_DOWN_ROWS, R >= 64 ? 4 : 1), 256, UL_DOWN_ROWS * ctx * sizeof(float), st>>>(
      (const half_t*)ehs, A, Z, M, ctx, R);
  if ((rc = check_launch("unet_lora_down_kernel"))) return rc;
  const unsigned gx = (unsigned)((KV / 8 + 127) / 128);
#define TB_UL_UP(RR, ROWS)                                                                                   \
  unet_lora_up_kernel<RR, ROWS><<<dim3(gx, (M + ROWS - 1) / ROWS), 128, 0, st>>>((half_t*)kv, Z, B, blk, M, KV, R, r, \
                                                                                 scaling)
  if (r <= 4) TB_UL_UP(4, 8);
  else if (r <= 8) TB_UL_UP(8, 4);
  else TB_UL_UP(16, 2);
#undef TB_UL_UP
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
#define TB_UL_DZ(RR, ROWS)                                                                                   \
  unet_lora_dz_kernel<RR, ROWS><<<dim3(n_adapters, (M + ROWS - 1) / ROWS), 128, 0, st>>>((const half_t*)dkv, B, off, dZ, \
                                                                                         M, KV, R, r, scaling)
  if (r <= 4) TB_UL_DZ(4, 8);
  else if (r <= 8) TB_UL_DZ(8, 4);
  else TB_UL_DZ(16, 2);
#undef TB_UL_DZ
  if ((rc = check_launch("unet_lora_dz_kernel"))) return rc;
  const int rows_per_cta = 16;
  const dim3 gb((KV / 8 + 127) / 128, (M + rows_per_cta - 1) / rows_per_cta);
  if (r <= 4)
    unet_lora_grad_b_kernel<4><<<gb, 128, 0, st>>>((const half_t*)dkv, Z, blk, dB, M, KV, R, r, scaling, rows_per_cta);
  else if (r <= 8)
    unet_lora_grad_b_kernel<8><<<gb, 128, 0, st>>>((const half_t*)dkv, Z, blk, dB, M, KV, R, r, scaling, rows_per_cta);
  else
    unet_lora_grad_b_kernel<UL_RMAX><<<gb, 128, 0, st>>>((const half_t*)dkv, Z, blk, dB, M, KV, R, r, scaling,
                                                         rows_per_cta);
  if ((rc = check_launch("unet_lora_grad_b_kernel"))) return rc;
  const int rows_a = 64, chunks_a = (M + rows_a - 1) / rows_a;
  unet_lora_grad_a_kernel<<<dim3((ctx + 255) / 256, ((R + 7) / 8) * chunks_a), 256, 0, st>>>(
      dZ, (const half_t*)ehs, dA, M, ctx, R, rows_a, chunks_a);
  if ((rc = check_launch("unet_lora_grad_a_kernel"))) return rc;
  unet_lora_dehs_kernel<<<dim3((ctx + 255) / 256, M), 256, 0, st>>>(dZ, A, d_ehs, ctx, R);
  return check_launch("unet_lora_dehs_kernel");
}
