// Flash attention forward / backward on tcgen05 + TMEM for the UNet self- and cross-attention
// (diffusers AttnProcessor2_0 -> F.scaled_dot_product_attention, no mask, no dropout).
//
// Layout: q/k/v/o are token-major fp16 [B, N, ld] with head h at columns [h*d, (h+1)*d); TMA reads a
// head's [128 x 64] box straight out of the fused projection output (4-D map: d, heads, N, B) and
// zero-fills the d..64 padding and the rows past N.  head_dim d in {40, 64, 80, 160} (any d % 8 == 0
// up to 192).  S / dP / dQ and the O / dK / dV accumulators live in TMEM.
//
// CTA = 192 threads: warps 0..3 = softmax / correction / epilogue (thread == tile row),
// warp 4 = TMA producer, warp 5 = MMA issuer.
#include "host_util.h"
#include "sm100.cuh"

namespace tb {

constexpr int ABOX = 16384;  // one [128 rows x 64 fp16] swizzled box

struct AttnParams {
  int Nq, Nk, heads, d, dn;  // dn = d rounded up to 16 (MMA N / K extent)
  float scale_log2;          // softmax scale * log2(e)
  float scale;               // softmax scale
  __half* O;                 // fwd: output; bwd: unused
  long long ldo;
  float* lse;                // [B, heads, Nq]   log2-domain logsumexp
  const float* delta;        // bwd: [B, heads, Nq] rowsum(dO * O)
  float* dQacc;              // bwd: fp32 [B, Nq, lddq] accumulated with red.add (may be null)
  long long lddq;
  __half* dK;                // bwd: [B, Nk, lddk] head h at h*d
  long long lddk;
  __half* dV;
  long long lddv;
  int n_inner;               // fwd: kv tiles; bwd: q tiles
};

__device__ __forceinline__ uint32_t sw128_off(int row, int chunk16) {
  // byte offset of 16-byte chunk `chunk16` (0..15 over 128 columns) of row `row` in a
  // [128 x 128] fp16 tile stored as two swizzled [128 x 64] boxes
  return (uint32_t)((chunk16 >> 3) * ABOX + row * 128 + (((chunk16 & 7) ^ (row & 7)) << 4));
}

__device__ __forceinline__ void named_bar_sync(int id, int n) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}

// ------------------------------------------------------------------------------------ forward
template <int NB, int STAGES>
__global__ void __launch_bounds__(192) attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                       const __grid_constant__ CUtensorMap tmK,
                                                       const __grid_constant__ CUtensorMap tmV,
                                                       const AttnParams p) {
  constexpr int TILE = NB * ABOX;
  constexpr int TMEM_COLS = NB == 3 ? 512 : 256;
  constexpr uint32_t S_COL = 0, O_COL = 128;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + TILE;
  uint8_t* sV = sK + STAGES * TILE;
  uint8_t* sP = sV + STAGES * TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * ABOX);
  uint64_t* bar_q = bars;
  uint64_t* bar_s = bars + 1;
  uint64_t* bar_p = bars + 2;
  uint64_t* bar_o = bars + 3;
  uint64_t* kv_full = bars + 4;
  uint64_t* kv_empty = bars + 4 + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 + 2 * STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;

  if (threadIdx.x == 160) {
    mbar_init(smem_u32(bar_q), 1);
    mbar_init(smem_u32(bar_s), 1);
    mbar_init(smem_u32(bar_p), 128);
    mbar_init(smem_u32(bar_o), 1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&kv_full[s]), 1);
      mbar_init(smem_u32(&kv_empty[s]), 1);
    }
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<TMEM_COLS>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      mbar_expect_tx(smem_u32(bar_q), TILE);
      for (int x = 0; x < NB; ++x) tma_load_4d(smem_u32(sQ + x * ABOX), &tmQ, smem_u32(bar_q), x * 64, h, q0, b);
      for (int j = 0; j < p.n_inner; ++j) {
        const int s = j % STAGES;
        const uint32_t ph = (j / STAGES) & 1;
        mbar_wait(smem_u32(&kv_empty[s]), ph ^ 1);
        const uint32_t fb = smem_u32(&kv_full[s]);
        mbar_expect_tx(fb, 2 * TILE);
        for (int x = 0; x < NB; ++x) {
          tma_load_4d(smem_u32(sK + s * TILE + x * ABOX), &tmK, fb, x * 64, h, j * 128, b);
          tma_load_4d(smem_u32(sV + s * TILE + x * ABOX), &tmV, fb, x * 64, h, j * 128, b);
        }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_f16(128, 128, 0, 0);
      const uint32_t idesc_o = umma_idesc_f16(128, p.dn, 0, 1);  // B = V is MN-major
      const int nks = p.dn / 16;
      mbar_wait(smem_u32(bar_q), 0);
      for (int j = 0; j < p.n_inner; ++j) {
        const int s = j % STAGES;
        const uint32_t ph = (j / STAGES) & 1;
        mbar_wait(smem_u32(&kv_full[s]), ph);
        tc_fence_after();
        const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK + s * TILE), aV = smem_u32(sV + s * TILE);
        for (int ks = 0; ks < nks; ++ks) {
          const uint32_t off = (ks >> 2) * ABOX + (ks & 3) * 32;
          umma_f16_ss(tmem + S_COL, umma_desc_sw128(aQ + off, 16, 1024),
                      umma_desc_sw128(aK + off, 16, 1024), idesc_s, ks > 0);
        }
        umma_commit(smem_u32(bar_s));
        mbar_wait(smem_u32(bar_p), j & 1);
        tc_fence_after();
        const uint32_t aP = smem_u32(sP);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t offp = (ks >> 2) * ABOX + (ks & 3) * 32;
          umma_f16_ss(tmem + O_COL, umma_desc_sw128(aP + offp, 16, 1024),
                      umma_desc_sw128(aV + ks * 2048, ABOX, 1024), idesc_o, (j > 0) || (ks > 0));
        }
        umma_commit(smem_u32(&kv_empty[s]));
        umma_commit(smem_u32(bar_o));
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax warps (row = thread)
    const int row = warp * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < p.n_inner; ++j) {
      const int kv_valid = p.Nk - j * 128;  // columns >= kv_valid are padding
      mbar_wait(smem_u32(bar_s), j & 1);
      tc_fence_after();
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(lane_addr + S_COL + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float v = (c * 32 + i < kv_valid) ? __uint_as_float(r[i]) : -INFINITY;
          mx = fmaxf(mx, v);
        }
      }
      const float m_new = fmaxf(m, mx * p.scale_log2);
      const float alpha = exp2f(m - m_new);
      if (j > 0) {
        mbar_wait(smem_u32(bar_o), (j - 1) & 1);  // PV_{j-1} retired: O readable, sP reusable
        tc_fence_after();
        if (__any_sync(0xffffffffu, alpha != 1.f)) {
          for (int c = 0; c < p.dn; c += 16) {
            uint32_t r[16];
            tmem_ld16(lane_addr + O_COL + c, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
            tmem_st16(lane_addr + O_COL + c, r);
          }
          tmem_st_wait();
        }
      }
      float rs = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(lane_addr + S_COL + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float pv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int col = c * 32 + g * 8 + i;
            const float e = exp2f(__uint_as_float(r[g * 8 + i]) * p.scale_log2 - m_new);
            pv[i] = col < kv_valid ? e : 0.f;
            rs += pv[i];
          }
          uint4 o;
          o.x = pack_half2(pv[0], pv[1]);
          o.y = pack_half2(pv[2], pv[3]);
          o.z = pack_half2(pv[4], pv[5]);
          o.w = pack_half2(pv[6], pv[7]);
          *reinterpret_cast<uint4*>(sP + sw128_off(row, c * 4 + g)) = o;
        }
      }
      l = l * alpha + rs;
      m = m_new;
      fence_async_smem();
      tc_fence_before();
      mbar_arrive(smem_u32(bar_p));
    }
    mbar_wait(smem_u32(bar_o), (p.n_inner - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.f / l;
    const int q = q0 + row;
    const bool ok = q < p.Nq;
    __half* op = p.O + ((long long)b * p.Nq + q) * p.ldo + h * p.d;
    for (int c = 0; c < p.dn; c += 16) {
      uint32_t r[16];
      tmem_ld16(lane_addr + O_COL + c, r);
      tmem_ld_wait();
      if (ok) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (c + g * 8 < p.d) {
            uint4 o;
            o.x = pack_half2(__uint_as_float(r[g * 8 + 0]) * inv_l, __uint_as_float(r[g * 8 + 1]) * inv_l);
            o.y = pack_half2(__uint_as_float(r[g * 8 + 2]) * inv_l, __uint_as_float(r[g * 8 + 3]) * inv_l);
            o.z = pack_half2(__uint_as_float(r[g * 8 + 4]) * inv_l, __uint_as_float(r[g * 8 + 5]) * inv_l);
            o.w = pack_half2(__uint_as_float(r[g * 8 + 6]) * inv_l, __uint_as_float(r[g * 8 + 7]) * inv_l);
            *reinterpret_cast<uint4*>(op + c + g * 8) = o;
          }
        }
      }
    }
    if (ok && p.lse) p.lse[((long long)b * p.heads + h) * p.Nq + q] = m + log2f(l);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<TMEM_COLS>(tmem);
}

// ------------------------------------------------------------------------------------ backward
// delta[b,h,q] = sum_c dO[b,q,h*d+c] * O[b,q,h*d+c]   (one warp per (b,q), lanes over heads*d/8 vectors)
__global__ void attn_delta_kernel(const __half* __restrict__ O, const __half* __restrict__ dO,
                                  long long ldo, long long lddo, float* __restrict__ delta, int B,
                                  int Nq, int heads, int d) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= (long long)B * Nq) return;
  const int b = row / Nq, q = row % Nq;
  const int vec_per_head = d / 8;
  for (int hh = 0; hh < heads; ++hh) {
    float acc = 0.f;
    for (int v = lane; v < vec_per_head; v += 32) {
      const uint4 a = *reinterpret_cast<const uint4*>(O + row * ldo + hh * d + v * 8);
      const uint4 g = *reinterpret_cast<const uint4*>(dO + row * lddo + hh * d + v * 8);
      const __half2* ah = reinterpret_cast<const __half2*>(&a);
      const __half2* gh = reinterpret_cast<const __half2*>(&g);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 x = __half22float2(ah[i]), y = __half22float2(gh[i]);
        acc += x.x * y.x + x.y * y.y;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) delta[((long long)b * heads + hh) * Nq + q] = acc;
  }
}

// CTA owns one 128-row KV tile of one (b, h) and loops over the Q tiles.
//   S^T = K Q^T ; P^T = exp2(S^T*c - L) ; dP^T = V dO^T ; dS^T = P^T o (dP^T - delta)
//   dV += P^T dO ; dK += dS^T Q ; dQ_i = dS K  (red.add into the fp32 dQ accumulator)
// TMEM: X = [0,128) is S^T, then dP^T, then dQ_i in turn; dV at 128, dK at 128+dn.
template <int NB, int STAGES>
__global__ void __launch_bounds__(192) attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                       const __grid_constant__ CUtensorMap tmK,
                                                       const __grid_constant__ CUtensorMap tmV,
                                                       const __grid_constant__ CUtensorMap tmdO,
                                                       const AttnParams p) {
  constexpr int TILE = NB * ABOX;
  constexpr bool INPLACE = NB == 3;  // dS^T overwrites P^T in place (shared memory budget)
  constexpr int TMEM_COLS = NB == 1 ? 256 : 512;
  constexpr uint32_t X_COL = 0, DV_COL = 128;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + TILE;
  uint8_t* sQ = sV + TILE;
  uint8_t* sdO = sQ + STAGES * TILE;
  uint8_t* sPT = sdO + STAGES * TILE;
  uint8_t* sDS = INPLACE ? sPT : sPT + 2 * ABOX;
  float* sL = reinterpret_cast<float*>(sDS + 2 * ABOX);  // [128] L_i, [128] delta_i
  float* sD = sL + 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sD + 128);
  uint64_t* bar_kv = bars;
  uint64_t* bar_s = bars + 1;
  uint64_t* bar_pt = bars + 2;
  uint64_t* bar_dp = bars + 3;
  uint64_t* bar_dv = bars + 4;
  uint64_t* bar_ds = bars + 5;
  uint64_t* bar_dq = bars + 6;
  uint64_t* bar_dqfree = bars + 7;
  uint64_t* q_full = bars + 8;
  uint64_t* q_empty = bars + 8 + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8 + 2 * STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kv0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const uint32_t DK_COL = DV_COL + p.dn;

  if (threadIdx.x == 160) {
    mbar_init(smem_u32(bar_kv), 1);
    mbar_init(smem_u32(bar_s), 1);
    mbar_init(smem_u32(bar_pt), 128);
    mbar_init(smem_u32(bar_dp), 1);
    mbar_init(smem_u32(bar_dv), 1);
    mbar_init(smem_u32(bar_ds), 128);
    mbar_init(smem_u32(bar_dq), 1);
    mbar_init(smem_u32(bar_dqfree), 128);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&q_full[s]), 1);
      mbar_init(smem_u32(&q_empty[s]), 1);
    }
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<TMEM_COLS>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      tma_prefetch_desc(&tmdO);
      mbar_expect_tx(smem_u32(bar_kv), 2 * TILE);
      for (int x = 0; x < NB; ++x) {
        tma_load_4d(smem_u32(sK + x * ABOX), &tmK, smem_u32(bar_kv), x * 64, h, kv0, b);
        tma_load_4d(smem_u32(sV + x * ABOX), &tmV, smem_u32(bar_kv), x * 64, h, kv0, b);
      }
      for (int i = 0; i < p.n_inner; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(smem_u32(&q_empty[s]), ph ^ 1);
        const uint32_t fb = smem_u32(&q_full[s]);
        mbar_expect_tx(fb, 2 * TILE);
        for (int x = 0; x < NB; ++x) {
          tma_load_4d(smem_u32(sQ + s * TILE + x * ABOX), &tmQ, fb, x * 64, h, i * 128, b);
          tma_load_4d(smem_u32(sdO + s * TILE + x * ABOX), &tmdO, fb, x * 64, h, i * 128, b);
        }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      const uint32_t idesc_kk = umma_idesc_f16(128, 128, 0, 0);     // S^T, dP^T : both K-major
      const uint32_t idesc_acc = umma_idesc_f16(128, p.dn, 0, 1);   // dV, dK : B MN-major
      const int nks = p.dn / 16;
      // dQ = dS K: A = dS^T buffer read MN-major, B = K MN-major; N split at 128 when dn > 128
      const int dq_n0 = p.dn > 128 ? 128 : p.dn;
      const int dq_n1 = p.dn - dq_n0;
      const uint32_t idesc_dq0 = umma_idesc_f16(128, dq_n0, 1, 1);
      const uint32_t idesc_dq1 = umma_idesc_f16(128, dq_n1 > 0 ? dq_n1 : 16, 1, 1);
      const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aPT = smem_u32(sPT), aDS = smem_u32(sDS);
      uint32_t ph_dqfree = 0;
      mbar_wait(smem_u32(bar_kv), 0);
      for (int i = 0; i < p.n_inner; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        const uint32_t aQ = smem_u32(sQ + s * TILE), adO = smem_u32(sdO + s * TILE);
        mbar_wait(smem_u32(&q_full[s]), ph);
        if (i > 0) {
          mbar_wait(smem_u32(bar_dqfree), ph_dqfree);
          ph_dqfree ^= 1;
        }
        tc_fence_after();
        for (int ks = 0; ks < nks; ++ks) {  // S^T = K Q^T
          const uint32_t off = (ks >> 2) * ABOX + (ks & 3) * 32;
          umma_f16_ss(tmem + X_COL, umma_desc_sw128(aK + off, 16, 1024),
                      umma_desc_sw128(aQ + off, 16, 1024), idesc_kk, ks > 0);
        }
        umma_commit(smem_u32(bar_s));
        mbar_wait(smem_u32(bar_pt), i & 1);  // P^T in smem, S^T drained from TMEM
        tc_fence_after();
        for (int ks = 0; ks < nks; ++ks) {  // dP^T = V dO^T
          const uint32_t off = (ks >> 2) * ABOX + (ks & 3) * 32;
          umma_f16_ss(tmem + X_COL, umma_desc_sw128(aV + off, 16, 1024),
                      umma_desc_sw128(adO + off, 16, 1024), idesc_kk, ks > 0);
        }
        umma_commit(smem_u32(bar_dp));
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {  // dV += P^T dO
          const uint32_t offp = (ks >> 2) * ABOX + (ks & 3) * 32;
          umma_f16_ss(tmem + DV_COL, umma_desc_sw128(aPT + offp, 16, 1024),
                      umma_desc_sw128(adO + ks * 2048, ABOX, 1024), idesc_acc, (i > 0) || (ks > 0));
        }
        if (INPLACE) umma_commit(smem_u32(bar_dv));
        mbar_wait(smem_u32(bar_ds), i & 1);  // dS^T in smem, dP^T drained
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {  // dK += dS^T Q
          const uint32_t offp = (ks >> 2) * ABOX + (ks & 3) * 32;
          umma_f16_ss(tmem + DK_COL, umma_desc_sw128(aDS + offp, 16, 1024),
                      umma_desc_sw128(aQ + ks * 2048, ABOX, 1024), idesc_acc, (i > 0) || (ks > 0));
        }
        if (p.dQacc) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {  // dQ_i = dS K   (A: [kv][q] read MN-major)
            umma_f16_ss(tmem + X_COL, umma_desc_sw128(aDS + ks * 2048, ABOX, 1024),
                        umma_desc_sw128(aK + ks * 2048, ABOX, 1024), idesc_dq0, ks > 0);
          }
        }
        umma_commit(smem_u32(bar_dq));
        if (p.dQacc && dq_n1 > 0) {
          // columns 128..dn of dQ reuse X once the first 128 have been drained
          mbar_wait(smem_u32(bar_dqfree), ph_dqfree);
          ph_dqfree ^= 1;
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            umma_f16_ss(tmem + X_COL, umma_desc_sw128(aDS + ks * 2048, ABOX, 1024),
                        umma_desc_sw128(aK + 2 * ABOX + ks * 2048, ABOX, 1024), idesc_dq1, ks > 0);
          }
          umma_commit(smem_u32(bar_dq));
        }
        umma_commit(smem_u32(&q_empty[s]));
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax / dS / epilogue
    const int row = warp * 32 + lane;  // kv row inside the tile (and q row for the dQ drain)
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    const bool kv_ok = kv0 + row < p.Nk;
    const long long bh = (long long)b * p.heads + h;
    uint32_t ph_dq = 0;
    for (int i = 0; i < p.n_inner; ++i) {
      const int q0 = i * 128;
      {
        const int q = q0 + row;
        sL[row] = q < p.Nq ? p.lse[bh * p.Nq + q] : INFINITY;
        sD[row] = q < p.Nq ? p.delta[bh * p.Nq + q] : 0.f;
      }
      named_bar_sync(1, 128);
      mbar_wait(smem_u32(bar_s), i & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(lane_addr + X_COL + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float pv[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int col = c * 32 + g * 8 + k;
            const float e = exp2f(__uint_as_float(r[g * 8 + k]) * p.scale_log2 - sL[col]);
            pv[k] = kv_ok ? e : 0.f;
          }
          uint4 o;
          o.x = pack_half2(pv[0], pv[1]);
          o.y = pack_half2(pv[2], pv[3]);
          o.z = pack_half2(pv[4], pv[5]);
          o.w = pack_half2(pv[6], pv[7]);
          *reinterpret_cast<uint4*>(sPT + sw128_off(row, c * 4 + g)) = o;
        }
      }
      fence_async_smem();
      tc_fence_before();
      mbar_arrive(smem_u32(bar_pt));

      mbar_wait(smem_u32(bar_dp), i & 1);
      if (INPLACE) mbar_wait(smem_u32(bar_dv), i & 1);  // dV MMA has finished reading P^T
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(lane_addr + X_COL + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint32_t off = sw128_off(row, c * 4 + g);
          const uint4 pq = *reinterpret_cast<const uint4*>(sPT + off);
          const __half2* ph2 = reinterpret_cast<const __half2*>(&pq);
          float ds[8];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 pf = __half22float2(ph2[k]);
            const int col = c * 32 + g * 8 + 2 * k;
            ds[2 * k] = pf.x * (__uint_as_float(r[g * 8 + 2 * k]) - sD[col]);
            ds[2 * k + 1] = pf.y * (__uint_as_float(r[g * 8 + 2 * k + 1]) - sD[col + 1]);
          }
          uint4 o;
          o.x = pack_half2(ds[0], ds[1]);
          o.y = pack_half2(ds[2], ds[3]);
          o.z = pack_half2(ds[4], ds[5]);
          o.w = pack_half2(ds[6], ds[7]);
          *reinterpret_cast<uint4*>(sDS + off) = o;
        }
      }
      fence_async_smem();
      tc_fence_before();
      mbar_arrive(smem_u32(bar_ds));

      const int n_chunks = (p.dQacc && p.dn > 128) ? 2 : 1;
      for (int ch = 0; ch < n_chunks; ++ch) {
        mbar_wait(smem_u32(bar_dq), ph_dq);
        ph_dq ^= 1;
        tc_fence_after();
        if (p.dQacc) {
          const int q = q0 + row;
          const bool q_ok = q < p.Nq;
          const int cbase = ch * 128;
          float* dq = p.dQacc + ((long long)b * p.Nq + q) * p.lddq + h * p.d + cbase;
          const int ncols = (p.dn - cbase) > 128 ? 128 : (p.dn - cbase);
          for (int c = 0; c < ncols; c += 16) {
            uint32_t r[16];
            tmem_ld16(lane_addr + X_COL + c, r);
            tmem_ld_wait();
            if (q_ok) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                if (cbase + c + g * 4 < p.d) {
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dq + c + g * 4),
                               "f"(__uint_as_float(r[g * 4 + 0]) * p.scale),
                               "f"(__uint_as_float(r[g * 4 + 1]) * p.scale),
                               "f"(__uint_as_float(r[g * 4 + 2]) * p.scale),
                               "f"(__uint_as_float(r[g * 4 + 3]) * p.scale)
                               : "memory");
                }
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(smem_u32(bar_dqfree));
      }
    }
    // ------------------------------------------------------------ dV / dK epilogue
    // (bar_dq of the last iteration covered every MMA issued by the CTA)
    for (int c = 0; c < p.dn; c += 16) {
      uint32_t rv[16], rk[16];
      tmem_ld16(lane_addr + DV_COL + c, rv);
      tmem_ld16(lane_addr + DK_COL + c, rk);
      tmem_ld_wait();
      if (kv_ok) {
        __half* dv = p.dV + ((long long)b * p.Nk + kv0 + row) * p.lddv + h * p.d;
        __half* dk = p.dK + ((long long)b * p.Nk + kv0 + row) * p.lddk + h * p.d;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (c + g * 8 < p.d) {
            uint4 o;
            o.x = pack_half2(__uint_as_float(rv[g * 8 + 0]), __uint_as_float(rv[g * 8 + 1]));
            o.y = pack_half2(__uint_as_float(rv[g * 8 + 2]), __uint_as_float(rv[g * 8 + 3]));
            o.z = pack_half2(__uint_as_float(rv[g * 8 + 4]), __uint_as_float(rv[g * 8 + 5]));
            o.w = pack_half2(__uint_as_float(rv[g * 8 + 6]), __uint_as_float(rv[g * 8 + 7]));
            *reinterpret_cast<uint4*>(dv + c + g * 8) = o;
            o.x = pack_half2(__uint_as_float(rk[g * 8 + 0]) * p.scale, __uint_as_float(rk[g * 8 + 1]) * p.scale);
            o.y = pack_half2(__uint_as_float(rk[g * 8 + 2]) * p.scale, __uint_as_float(rk[g * 8 + 3]) * p.scale);
            o.z = pack_half2(__uint_as_float(rk[g * 8 + 4]) * p.scale, __uint_as_float(rk[g * 8 + 5]) * p.scale);
            o.w = pack_half2(__uint_as_float(rk[g * 8 + 6]) * p.scale, __uint_as_float(rk[g * 8 + 7]) * p.scale);
            *reinterpret_cast<uint4*>(dk + c + g * 8) = o;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<TMEM_COLS>(tmem);
}

}  // namespace tb

// ------------------------------------------------------------------------------------ host side
using namespace tb;

static int make_head_map(CUtensorMap* m, const void* base, long long ld, int B, int heads, int N,
                         int d) {
  uint64_t dims[4] = {(uint64_t)d, (uint64_t)heads, (uint64_t)N, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)d * 2, (uint64_t)ld * 2, (uint64_t)N * ld * 2};
  uint32_t box[4] = {64, 1, 128, 1};
  return make_tmap_f16(m, base, 4, dims, strides, box);
}

static int check_attn_args(const char* fn, int B, int heads, int Nq, int Nk, int d, long long ldq,
                           long long ldk, long long ldv) {
  TB_REQUIRE(B > 0 && heads > 0 && Nq > 0 && Nk > 0, TB_E_SHAPE, "%s: bad sizes", fn);
  TB_REQUIRE(d % 8 == 0 && d >= 8 && d <= 192, TB_E_SHAPE, "%s: head_dim %d unsupported (d %% 8 == 0, d <= 192)", fn, d);
  TB_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0, TB_E_ALIGN, "%s: row strides must be multiples of 8", fn);
  TB_REQUIRE(ldq >= (long long)heads * d && ldk >= (long long)heads * d && ldv >= (long long)heads * d,
             TB_E_SHAPE, "%s: row stride smaller than heads*d", fn);
  return TB_OK;
}

template <int NB, int STAGES>
static int launch_attn_fwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                           const AttnParams& p, int B, cudaStream_t st) {
  constexpr int smem = NB * ABOX * (1 + 2 * STAGES) + 2 * ABOX + 256 + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel<NB, STAGES>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(attn_fwd<%d,%d>, %d): %s", NB, STAGES, smem, cudaGetErrorString(e));
      return TB_E_CUDA;
    }
    configured = true;
  }
  dim3 grid((p.Nq + 127) / 128, p.heads, B);
  attn_fwd_kernel<NB, STAGES><<<grid, 192, smem, st>>>(tq, tk, tv, p);
  return check_launch("attn_fwd_kernel");
}

template <int NB, int STAGES>
static int launch_attn_bwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                           const CUtensorMap& tdo, const AttnParams& p, int B, cudaStream_t st) {
  constexpr int smem = NB * ABOX * (2 + 2 * STAGES) + (NB == 3 ? 2 : 4) * ABOX + 1024 + 256 + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel<NB, STAGES>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(attn_bwd<%d,%d>, %d): %s", NB, STAGES, smem, cudaGetErrorString(e));
      return TB_E_CUDA;
    }
    configured = true;
  }
  dim3 grid((p.Nk + 127) / 128, p.heads, B);
  attn_bwd_kernel<NB, STAGES><<<grid, 192, smem, st>>>(tq, tk, tv, tdo, p);
  return check_launch("attn_bwd_kernel");
}

extern "C" int tb_attn_fwd_f16(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                               int64_t ldv, void* o, int64_t ldo, float* lse, int B, int heads, int Nq,
                               int Nk, int d, float scale, void* stream) {
  int rc = tb_check_device();
  if (rc) return rc;
  TB_REQUIRE(q && k && v && o, TB_E_ARG, "tb_attn_fwd_f16: null pointer");
  rc = check_attn_args("tb_attn_fwd_f16", B, heads, Nq, Nk, d, ldq, ldk, ldv);
  if (rc) return rc;
  TB_REQUIRE(ldo % 8 == 0, TB_E_ALIGN, "tb_attn_fwd_f16: ldo %% 8");
  CUtensorMap tq, tk, tv;
  if ((rc = make_head_map(&tq, q, ldq, B, heads, Nq, d))) return rc;
  if ((rc = make_head_map(&tk, k, ldk, B, heads, Nk, d))) return rc;
  if ((rc = make_head_map(&tv, v, ldv, B, heads, Nk, d))) return rc;
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.Nq = Nq; p.Nk = Nk; p.heads = heads; p.d = d; p.dn = (d + 15) / 16 * 16;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.O = (__half*)o; p.ldo = ldo; p.lse = lse;
  p.n_inner = (Nk + 127) / 128;
  const int nb = (d + 63) / 64;
  cudaStream_t st = (cudaStream_t)stream;
  if (nb == 1) return launch_attn_fwd<1, 2>(tq, tk, tv, p, B, st);
  if (nb == 2) return launch_attn_fwd<2, 2>(tq, tk, tv, p, B, st);
  return launch_attn_fwd<3, 1>(tq, tk, tv, p, B, st);
}

extern "C" int tb_attn_bwd_f16(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                               int64_t ldv, const void* o, int64_t ldo, const void* dO, int64_t lddo,
                               const float* lse, float* delta, float* dQacc, int64_t lddq, void* dK,
                               int64_t lddk, void* dV, int64_t lddv, int B, int heads, int Nq, int Nk,
                               int d, float scale, void* stream) {
  int rc = tb_check_device();
  if (rc) return rc;
  TB_REQUIRE(q && k && v && o && dO && lse && delta && dK && dV, TB_E_ARG, "tb_attn_bwd_f16: null pointer");
  rc = check_attn_args("tb_attn_bwd_f16", B, heads, Nq, Nk, d, ldq, ldk, ldv);
  if (rc) return rc;
  TB_REQUIRE(ldo % 8 == 0 && lddo % 8 == 0 && lddk % 8 == 0 && lddv % 8 == 0 && lddq % 4 == 0,
             TB_E_ALIGN, "tb_attn_bwd_f16: stride alignment");
  cudaStream_t st = (cudaStream_t)stream;
  {
    const long long rows = (long long)B * Nq;
    const int wpb = 8;
    attn_delta_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, st>>>(
        (const __half*)o, (const __half*)dO, ldo, lddo, delta, B, Nq, heads, d);
    if ((rc = check_launch("attn_delta_kernel"))) return rc;
  }
  if (dQacc) {
    cudaError_t e = cudaMemsetAsync(dQacc, 0, (size_t)B * Nq * lddq * sizeof(float), st);
    TB_REQUIRE(e == cudaSuccess, TB_E_CUDA, "tb_attn_bwd_f16: memset dQacc: %s", cudaGetErrorString(e));
  }
  CUtensorMap tq, tk, tv, tdo;
  if ((rc = make_head_map(&tq, q, ldq, B, heads, Nq, d))) return rc;
  if ((rc = make_head_map(&tk, k, ldk, B, heads, Nk, d))) return rc;
  if ((rc = make_head_map(&tv, v, ldv, B, heads, Nk, d))) return rc;
  if ((rc = make_head_map(&tdo, dO, lddo, B, heads, Nq, d))) return rc;
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.Nq = Nq; p.Nk = Nk; p.heads = heads; p.d = d; p.dn = (d + 15) / 16 * 16;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.lse = const_cast<float*>(lse);
  p.delta = delta;
  p.dQacc = dQacc; p.lddq = lddq;
  p.dK = (__half*)dK; p.lddk = lddk;
  p.dV = (__half*)dV; p.lddv = lddv;
  p.n_inner = (Nq + 127) / 128;
  const int nb = (d + 63) / 64;
  if (nb == 1) return launch_attn_bwd<1, 2>(tq, tk, tv, tdo, p, B, st);
  if (nb == 2) return launch_attn_bwd<2, 1>(tq, tk, tv, tdo, p, B, st);
  return launch_attn_bwd<3, 1>(tq, tk, tv, tdo, p, B, st);
}
