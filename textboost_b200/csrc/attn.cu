// Flash attention forward / backward on tcgen05 + TMEM for the UNet self- and cross-attention
// (diffusers AttnProcessor2_0 -> F.scaled_dot_product_attention, no mask, no dropout) and, with
// causal=1, for the CLIP text encoder (transformers CLIPAttention, 77 tokens, head_dim 64).
//
// Layout: q/k/v/o are token-major fp16 [B, N, ld] with head h at columns [h*d, (h+1)*d); TMA reads a
// head's [128 x 64] box straight out of the fused projection output (4-D map: d, heads, N, B) and
// zero-fills the d..64 padding and the rows past N.  head_dim d % 8 == 0, d <= 192.
//
// The softmax is the bound, not the tensor core: a 128x128 score tile costs 16384 exponentials = 1024
// MUFU cycles per SM against 384 tensor cycles (d = 40).  So the kernels are organised around the
// per-element instruction count of the softmax threads:
//   * scores are read from TMEM in two passes (row max, then exp) so no 128-register row is held;
//   * exp2 runs as ex2.approx.f16x2 on the packed pair (the result IS the fp16 P the MMA consumes);
//   * forward: P goes back to TMEM (aliasing the first 64 score columns) and the PV product is a
//     TMEM-A MMA, so no shared-memory round trip; the row sum is a second tiny MMA against a tile of
//     ones (fp32 accumulation of exactly the P that multiplies V); the running max is only refreshed
//     when it grows by more than 2^8, so O is rarely rescaled;
//   * two CTAs co-reside per SM, one's MMAs overlapping the other's softmax.
//
// CTA = 320 threads: warps 0..7 = softmax / correction / epilogue (two threads per tile row: warp w owns
// TMEM lane quarter w & 3 and the 64-column half w >> 2), warp 8 = TMA producer, warp 9 = MMA issuer.
// Sixteen softmax warps per SM (two CTAs) are what it takes to keep the MUFU pipe busy: with one thread
// per row the per-thread TMEM-load / exp latency chain leaves it ~40 % utilised (ncu, profiles/).
#include <stdlib.h>

#include <type_traits>

#include "host_util.h"
#include "sm100.cuh"

namespace tb {

constexpr int ABOX = 16384;  // one [128 rows x 64 fp16] swizzled box

struct AttnParams {
  int Nq, Nk, heads, d, dn;  // dn = d rounded up to 16 (MMA N / K extent)
  int causal;                // key j visible to query i iff j <= i
  int tmem_cols;             // power of two >= the columns the kernel uses
  float scale_log2;          // softmax scale * log2(e)
  float scale;               // softmax scale
  tb::half_t* O;                 // fwd: output; bwd: unused
  long long ldo;
  float* lse;                // [B, heads, Nq]   log2-domain logsumexp
  const float* delta;        // bwd: [B, heads, Nq] rowsum(dO * O)
  float* dQacc;              // bwd: fp32 [B, Nq, lddq] accumulated with red.add (may be null)
  long long lddq;
  tb::half_t* dQ16;              // bwd, single KV tile (Nk <= 128): dQ written once as fp16 [B, Nq, lddq16] (no accumulator)
  long long lddq16;
  tb::half_t* dK;                // bwd: [B, Nk, lddk] head h at h*d
  long long lddk;
  tb::half_t* dV;
  long long lddv;
  int n_inner;               // fwd: kv tiles; bwd: q tiles
  int qsplit;                // bwd: CTAs per KV tile, each owning a contiguous range of Q tiles (1 = whole loop)
  float* dkv_ws;             // bwd, qsplit > 1: fp32 [2][B][Nk][heads*d] accumulators (dV, dK) filled with red.add
  int experiment;            // timing experiments only (TB_ATTN_EXPERIMENT): 1 = skip the dQ reds
  long long* trace;          // debug: clock64 timestamps of CTA (0,0,0), [iter][event][warp]; normally null
};

__device__ __forceinline__ uint32_t sw128_off(int row, int chunk16) {
  // byte offset of 16-byte chunk `chunk16` (0..15 over 128 columns) of row `row` in a
  // [128 x 128] fp16 tile stored as two swizzled [128 x 64] boxes
  return (uint32_t)((chunk16 >> 3) * ABOX + row * 128 + (((chunk16 & 7) ^ (row & 7)) << 4));
}

// debug timeline (tb_attn_debug_trace): event e of iteration j seen by warp w of CTA (0,0,0)
#define TB_TRACE(j, e)                                                                              \
  if (p.trace && lane == 0 && (j) < 16 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)    \
  p.trace[((j) * 8 + (e)) * 10 + warp] = clock64()

// the same for the 19-warp backward kernel: [iter][event][warp] with 20 warp slots
#define TB_TRACE3(j, e)                                                                             \
  if (p.trace && lane == 0 && (j) < 16 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)    \
  p.trace[((j) * 8 + (e)) * 20 + warp] = clock64()

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
// exp2 of two fp32 exponents, returned as the packed fp16 pair {lo, hi}.
// Two MUFU.EX2 (fp32) + one F2FP.  (ex2.approx.f16x2 also lowers to two MUFU ops, MUFU.EX2.F16, but those
// issue at roughly half the fp32 rate on sm_100: measured 0.59 ms vs the fp32 form on the 64x64 layer.)
__device__ __forceinline__ uint32_t exp2_pack(float lo, float hi) {
#ifdef TB_ATTN_EX2_F16X2
  uint32_t p, e;
  asm("cvt.rn." TB_H16X2 ".f32 %0, %1, %2;" : "=r"(p) : "f"(hi), "f"(lo));
#ifdef TB_BF16
  asm("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(e) : "r"(p));
#else
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(e) : "r"(p));
#endif
  return e;
#else
  float a, b;
  uint32_t e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(a) : "f"(lo));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(b) : "f"(hi));
  asm("cvt.rn." TB_H16X2 ".f32 %0, %1, %2;" : "=r"(e) : "f"(b), "f"(a));
  return e;
#endif
}
// 2^x on the FMA pipe (no MUFU): x = r + f with r = rint(x) taken from the mantissa of x + 1.5*2^23 and f in
// [-0.5, 0.5]; 2^f by a degree-4 minimax polynomial (max relative error 2.7e-6, far below the fp16 rounding of
// P), 2^r added into the exponent field.  The softmax phases are MUFU-bound (16 ex2/clk/SM: 1024 cycles per
// 128x128 tile, measured ~1830 for the P phase), while the FMA pipe issues 128/clk: moving a share of the
// exponentials here shortens the phase (the same split FlashAttention-4 uses).  Inputs are <= ~0 here; anything
// below -126 (masked / padded entries arrive as -inf) returns ~2^-126, which rounds to 0 in fp16.
// Measured on the v2 backward (64x64 level): 25 % of the exponentials on the FMA pipe leaves the kernel time
// unchanged (1197 vs 1205 us) -- the MUFU cycles saved are paid back in issue slots (10 instructions per
// polynomial exponential) -- so the default keeps every exponential on MUFU; the switch stays for experiments.
#ifndef TB_ATTN_POLY_PACKS
#define TB_ATTN_POLY_PACKS 0  // of every 4 packed pairs (8 exponentials), how many go to the FMA pipe
#endif
#ifndef TB_ATTN_FWD_POLY_PACKS
#define TB_ATTN_FWD_POLY_PACKS 0  // the same switch for the forward kernel
#endif
__device__ __forceinline__ float exp2_poly(float x) {
  x = fmaxf(x, -126.f);
  const float xf = x + 12582912.f;
  const float r = xf - 12582912.f;
  const float f = x - r;
  float q = 0.009570101276040077f;
  q = fmaf(q, f, 0.05591785907745361f);
  q = fmaf(q, f, 0.240247443318367f);
  q = fmaf(q, f, 0.6931217908859253f);
  q = fmaf(q, f, 0.9999992847442627f);
  return __int_as_float(__float_as_int(q) + (__float_as_int(xf) << 23));
}
__device__ __forceinline__ uint32_t exp2_pack_poly(float lo, float hi) {
  uint32_t e;
  const float a = exp2_poly(lo), b = exp2_poly(hi);
  asm("cvt.rn." TB_H16X2 ".f32 %0, %1, %2;" : "=r"(e) : "f"(b), "f"(a));
  return e;
}
// pack k (0..3) of a group of four: the last TB_ATTN_POLY_PACKS packs use the polynomial
template <int K>
__device__ __forceinline__ uint32_t exp2_pack_mix(float lo, float hi) {
  if (K >= 4 - TB_ATTN_POLY_PACKS) return exp2_pack_poly(lo, hi);
  return exp2_pack(lo, hi);
}
__device__ __forceinline__ uint32_t hmul2_u32(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("mul.rn." TB_H16X2 " %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ void tmem_alloc_rt(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_rt(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
      "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
      "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

// ------------------------------------------------------------------------------------ forward
// TMEM: S at [0,128) (fp32 scores; P, packed fp16, overwrites [0,64) once the scores are consumed),
//       O at [128, 128+dn), row sum L at [128+dn, 128+dn+16).
template <int NB, int STAGES>
__global__ void __launch_bounds__(320, NB == 1 ? 2 : 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  constexpr int TILE = NB * ABOX;
  constexpr uint32_t S_COL = 0, P_COL = 0, O_COL = 128;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + TILE;
  uint8_t* sV = sK + STAGES * TILE;
  uint8_t* sOnes = sV + STAGES * TILE;
  float* sMax = reinterpret_cast<float*>(sOnes + ABOX);  // [2][128] partial row maxima of the two halves
  uint64_t* bars = reinterpret_cast<uint64_t*>(sMax + 256);
  uint64_t* bar_q = bars;
  uint64_t* bar_s = bars + 1;
  uint64_t* bar_p = bars + 2;
  uint64_t* bar_o = bars + 3;
  uint64_t* kv_full = bars + 4;
  uint64_t* kv_empty = bars + 4 + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 + 2 * STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const uint32_t L_COL = O_COL + p.dn;
  // causal: kv tiles entirely above the diagonal contribute nothing
  const int n_kv = p.causal ? min(p.n_inner, (min(q0 + 128, p.Nq) + 127) / 128) : p.n_inner;

  if (threadIdx.x == 288) {
    mbar_init(smem_u32(bar_q), 1);
    mbar_init(smem_u32(bar_s), 1);
    mbar_init(smem_u32(bar_p), 256);
    mbar_init(smem_u32(bar_o), 1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&kv_full[s]), 1);
      mbar_init(smem_u32(&kv_empty[s]), 1);
    }
    mbar_fence_init();
  }
  {  // tile of ones (B operand of the row-sum MMA); every element equal, so the swizzle is irrelevant
    const uint4 ones = make_uint4(TB_ONE_X2, TB_ONE_X2, TB_ONE_X2, TB_ONE_X2);  // packed pair of 1.0
    for (int i = threadIdx.x; i < ABOX / 16; i += 320) reinterpret_cast<uint4*>(sOnes)[i] = ones;
    fence_async_smem();
  }
  if (warp == 0) tmem_alloc_rt(smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 8) {
    // ---------------------------------------------------------------- TMA producer (warp-uniform)
    if (elect_one()) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      mbar_expect_tx(smem_u32(bar_q), TILE);
      for (int x = 0; x < NB; ++x) tma_load_4d(smem_u32(sQ + x * ABOX), &tmQ, smem_u32(bar_q), x * 64, h, q0, b);
    }
    __syncwarp();
    for (int j = 0; j < n_kv; ++j) {
      const int s = j % STAGES;
      const uint32_t ph = (j / STAGES) & 1;
      mbar_wait(smem_u32(&kv_empty[s]), ph ^ 1);
      if (elect_one()) {
        const uint32_t fb = smem_u32(&kv_full[s]);
        mbar_expect_tx(fb, 2 * TILE);
        for (int x = 0; x < NB; ++x) {
          tma_load_4d(smem_u32(sK + s * TILE + x * ABOX), &tmK, fb, x * 64, h, j * 128, b);
          tma_load_4d(smem_u32(sV + s * TILE + x * ABOX), &tmV, fb, x * 64, h, j * 128, b);
        }
      }
      __syncwarp();
    }
  } else if (warp == 9) {
    // ---------------------------------------------------------------- MMA issuer (warp-uniform)
    const uint32_t idesc_s = umma_idesc_f16(128, 128, 0, 0);
    const uint32_t idesc_o = umma_idesc_f16(128, p.dn, 0, 1);  // B = V is MN-major
    const uint32_t idesc_l = umma_idesc_f16(128, 16, 0, 1);
    const int nks = p.dn / 16;
    const uint32_t hi = umma_desc_hi_sw128(1024);
    const uint32_t q_lo = umma_desc_lo(smem_u32(sQ), 16);
    const uint32_t ones_lo = umma_desc_lo(smem_u32(sOnes), ABOX);
    mbar_wait(smem_u32(bar_q), 0);
    for (int j = 0; j < n_kv; ++j) {
      const int s = j % STAGES;
      const uint32_t ph = (j / STAGES) & 1;
      const uint32_t k_lo = umma_desc_lo(smem_u32(sK + s * TILE), 16);
      const uint32_t v_lo = umma_desc_lo(smem_u32(sV + s * TILE), ABOX);
      mbar_wait(smem_u32(&kv_full[s]), ph);
      tc_fence_after();
      if (elect_one()) {
        for (int ks = 0; ks < nks; ++ks) {
          const uint32_t off = ((ks >> 2) * ABOX + (ks & 3) * 32) >> 4;
          umma_f16_ss(tmem + S_COL, umma_desc_pack(q_lo + off, hi), umma_desc_pack(k_lo + off, hi), idesc_s, ks > 0);
        }
        umma_commit(smem_u32(bar_s));  // also: every earlier MMA (PV of tile j-1) has retired
      }
      __syncwarp();
      TB_TRACE(j, 0);
      mbar_wait(smem_u32(bar_p), j & 1);
      tc_fence_after();
      TB_TRACE(j, 1);
      // keys past Nk have P == 0: skip their k-steps
      const int kvalid = min(128, p.Nk - j * 128);
      const int nkk = (kvalid + 15) >> 4;
      if (elect_one()) {
        for (int ks = 0; ks < nkk; ++ks) {
          umma_f16_ts(tmem + O_COL, tmem + P_COL + ks * 8, umma_desc_pack(v_lo + ks * 128, hi), idesc_o,
                      (j > 0) || (ks > 0));
          umma_f16_ts(tmem + L_COL, tmem + P_COL + ks * 8, umma_desc_pack(ones_lo + ks * 128, hi), idesc_l,
                      (j > 0) || (ks > 0));
        }
        umma_commit(smem_u32(&kv_empty[s]));
        if (j == n_kv - 1) umma_commit(smem_u32(bar_o));
      }
      __syncwarp();
      TB_TRACE(j, 2);
    }
  } else {
    // ---------------------------------------------------------------- softmax warps (2 threads per row)
    const int quarter = warp & 3, half = warp >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    const int cbase = half * 64;  // first score column of this thread
    float m_run = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      // columns >= limit are masked (padding keys, and keys above the causal diagonal)
      int limit = p.Nk - j * 128;
      if (p.causal) limit = min(limit, q0 + row + 1 - j * 128);
      const bool masked = __any_sync(0xffffffffu, limit < cbase + 64);
      TB_TRACE(j, 3);
      mbar_wait(smem_u32(bar_s), j & 1);
      tc_fence_after();
      TB_TRACE(j, 4);
      uint32_t r[64];
      tmem_ld32(lane_addr + S_COL + cbase, r);
      tmem_ld32(lane_addr + S_COL + cbase + 32, r + 32);
      tmem_ld_wait();
      TB_TRACE(j, 5);
      if (masked) {
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (cbase + i >= limit) r[i] = 0xff800000u;  // -inf
      }
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < 64; i += 2) mx = fmax3(mx, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
      sMax[half * 128 + row] = mx;
      // all 256 threads have their scores in registers past this barrier: P may now overwrite S
      named_bar_sync(1, 256);
      TB_TRACE(j, 6);
      mx = fmaxf(mx, sMax[(half ^ 1) * 128 + row]);
      const float m_tile = mx * p.scale_log2;
      if (j == 0) {
        m_run = m_tile;
      } else {
        // lazy rescale: keep the old reference max unless the row max grew by more than 2^8
        const bool need = m_tile > m_run + 8.f;
        if (__any_sync(0xffffffffu, need)) {
          const float alpha = need ? exp2f(m_run - m_tile) : 1.f;
          // O and the row sum L are contiguous: the two halves split the 16-column chunks
          for (int c = half * 16; c < p.dn + 16; c += 32) {
            uint32_t t[16];
            tmem_ld16(lane_addr + O_COL + c, t);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * alpha);
            tmem_st16(lane_addr + O_COL + c, t);
          }
        }
        if (need) m_run = m_tile;
      }
      // P = exp2(S*c - m) as packed fp16, written over the consumed score columns [0, 64)
      const float neg_m = -m_run;
      uint32_t pk[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float e0 = fmaf(__uint_as_float(r[2 * i]), p.scale_log2, neg_m);
        const float e1 = fmaf(__uint_as_float(r[2 * i + 1]), p.scale_log2, neg_m);
        pk[i] = (i & 3) >= 4 - TB_ATTN_FWD_POLY_PACKS ? exp2_pack_poly(e0, e1) : exp2_pack(e0, e1);
      }
      tmem_st32(lane_addr + P_COL + half * 32, pk);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(bar_p));
      TB_TRACE(j, 7);
    }
    mbar_wait(smem_u32(bar_o), 0);
    tc_fence_after();
    float l;
    {
      uint32_t t[8];
      tmem_ld8(lane_addr + L_COL, t);
      tmem_ld_wait();
      l = __uint_as_float(t[0]);
    }
    const float inv_l = 1.f / l;
    const int q = q0 + row;
    const bool ok = q < p.Nq;
    tb::half_t* op = p.O + ((long long)b * p.Nq + q) * p.ldo + h * p.d;
    for (int c = half * 16; c < p.dn; c += 32) {
      uint32_t t[16];
      tmem_ld16(lane_addr + O_COL + c, t);
      tmem_ld_wait();
      if (ok) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (c + g * 8 < p.d) {
            uint4 o;
            o.x = pack_half2(__uint_as_float(t[g * 8 + 0]) * inv_l, __uint_as_float(t[g * 8 + 1]) * inv_l);
            o.y = pack_half2(__uint_as_float(t[g * 8 + 2]) * inv_l, __uint_as_float(t[g * 8 + 3]) * inv_l);
            o.z = pack_half2(__uint_as_float(t[g * 8 + 4]) * inv_l, __uint_as_float(t[g * 8 + 5]) * inv_l);
            o.w = pack_half2(__uint_as_float(t[g * 8 + 6]) * inv_l, __uint_as_float(t[g * 8 + 7]) * inv_l);
            *reinterpret_cast<uint4*>(op + c + g * 8) = o;
          }
        }
      }
    }
    if (ok && half == 0 && p.lse) p.lse[((long long)b * p.heads + h) * p.Nq + q] = m_run + log2f(l);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc_rt(tmem, (uint32_t)p.tmem_cols);
}

// ------------------------------------------------------------------------------------ forward, v2
// Long, non-causal sequences with head_dim <= 128 (every UNet self-attention level that matters: 64x64 at d = 40,
// 32x32 at d = 80, and head_dim 64 of SD-2.x).  attn_fwd_kernel serialises  S MMA -> softmax -> PV MMA  inside a
// CTA and leans on a second resident CTA for overlap; measured 1756 cycles per 128x128 tile per SM against the
// 1024-cycle exponential bound (and 1.5x behind cuDNN's fused kernel).  Here ONE CTA per SM owns TWO query tiles
// (256 rows) of a (b, h) and ping-pongs them through the tensor core:
//   * warps 0-3 = softmax of tile 0, warps 4-7 = softmax of tile 1, ONE THREAD PER ROW: the row maximum and the row
//     sum need no exchange between threads, hence no CTA-level barrier anywhere in the loop; warp 8 = TMA (K/V
//     ring), warp 9 = MMA issuer;
//   * head_dim <= 64 (SPLIT): TMEM = S0 [0,128) S1 [128,256) P0 [256,320) P1 [320,384) O0 [384,448) O1 [448,512).
//     P has its own columns, so S_t(j+1) = Q_t K(j+1)^T is issued as soon as the softmax warps have READ S_t(j)
//     (bar_sfree), a whole exp pass before they need it: the softmax warps never wait for the tensor core, the two
//     tiles only compete for the MUFU.  (Measured with P aliasing S, i.e. S_t(j+1) behind PV_t(j): 1230 cycles from
//     "P published" to "next S seen" per tile -- wake-up + 19 MMAs + commit latency -- of a 3170-cycle iteration.)
//   * head_dim 65..128: P_t overwrites the first 64 columns of S_t (no room for separate P columns next to two
//     O accumulators), issue order  [P0(j)] PV0(j), S0(j+1)  [P1(j)] PV1(j), S1(j+1);
//     tcgen05.mma executes in issue order, so S_t(j+1) complete implies PV_t(j) complete (O_t may be rescaled);
//   * the scores are read from TMEM in two passes of four 32-column chunks (max, then exp) so a thread never holds
//     more than two chunks (TMEM reads run at ~700 B/clk/SM, scripts/micro/softmax_pipes.cu); P leaves chunk by
//     chunk (16 packed columns);
//   * scale-and-subtract runs as packed fp32x2 FMAs (FFMA2: one issue slot per two scores), the exponentials as
//     MUFU.EX2 and -- TB_ATTN_FWD2_POLY of every 8 packed pairs -- as a degree-3 polynomial on the FMA pipe
//     (Cody-Waite split + exponent insertion, packed): the loop is bound by the 16 ex2/clk/SM of the MUFU, so
//     moving a share of the exponentials to the FMA pipe shortens it (FlashAttention-4's split);
//   * the row sum is accumulated in fp32 registers from the unrounded exponentials (as FlashAttention does): the
//     ones-tile MMA of attn_fwd_kernel costs 8 extra tcgen05.mma per tile at ~14 cycles each.
#ifndef TB_ATTN_FWD2_POLY
#define TB_ATTN_FWD2_POLY 2
#endif
// The two softmax warp groups have identical loops; started together they run in lockstep (both in the max pass,
// then both in the exp pass) and the MUFU idles during the max pass.  Tile 1 starts TB_ATTN_FWD2_SKEW cycles late, so
// that its max pass falls into the exp pass of tile 0 and vice versa (the offset persists: the loops are symmetric).
#ifndef TB_ATTN_FWD2_SKEW
#define TB_ATTN_FWD2_SKEW 1000
#endif
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t sub_f32x2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// two exponentials 2^e0, 2^e1 (e <= ~8) on MUFU
__device__ __forceinline__ void exp2_pair_mufu(uint64_t e, float& p0, float& p1) {
  float e0, e1;
  unpack_f32x2(e, e0, e1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(e0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(e1));
}
// the same on the FMA / ALU pipes: x = r + f, r = rint(x) from the magic-number add, 2^f by a degree-3 minimax
// polynomial on [-0.5, 0.5] (max relative error 1.0e-4, a fifth of an fp16 ulp of P), 2^r into the exponent field.
__device__ __forceinline__ void exp2_pair_poly(uint64_t e, float& p0, float& p1) {
  float e0, e1;
  unpack_f32x2(e, e0, e1);
  e = pack_f32x2(fmaxf(e0, -125.f), fmaxf(e1, -125.f));
  const uint64_t magic = pack_f32x2(12582912.f, 12582912.f);
  const uint64_t xf = add_f32x2(e, magic);
  const uint64_t fr = sub_f32x2(e, sub_f32x2(xf, magic));
  uint64_t q = pack_f32x2(0.05550410866f, 0.05550410866f);
  q = fma_f32x2(q, fr, pack_f32x2(0.2402265069f, 0.2402265069f));
  q = fma_f32x2(q, fr, pack_f32x2(0.6931471806f, 0.6931471806f));
  q = fma_f32x2(q, fr, pack_f32x2(1.0f, 1.0f));
  float q0, q1, x0, x1;
  unpack_f32x2(q, q0, q1);
  unpack_f32x2(xf, x0, x1);
  p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(x0) << 23));
  p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(x1) << 23));
}

template <int NB, int STAGES>
__global__ void __launch_bounds__(384, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  constexpr int TILE = NB * ABOX;
  constexpr bool SPLIT = NB == 1;  // P in its own TMEM columns
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;                  // two query tiles
  uint8_t* sK = sQ + 2 * TILE;
  uint8_t* sV = sK + STAGES * TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + STAGES * TILE);
  uint64_t* bar_q = bars;              // [2]  Q_t landed
  uint64_t* bar_s = bars + 2;          // [2]  S_t(j) complete                      (MMA -> softmax)
  uint64_t* bar_p = bars + 4;          // [2]  P_t(j) stored                        (softmax -> MMA)
  uint64_t* bar_o = bars + 6;          // [2]  last PV_t complete
  uint64_t* bar_sfree = bars + 8;      // [2]  S_t(j) is in registers               (softmax -> MMA, SPLIT)
  uint64_t* bar_pfree = bars + 10;     // [2]  PV_t(j) complete: P_t / O_t are free (MMA -> softmax, SPLIT)
  uint64_t* kv_full = bars + 12;
  uint64_t* kv_empty = bars + 12 + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12 + 2 * STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256, h = blockIdx.y, b = blockIdx.z;
  const int n_kv = p.n_inner;
  const bool two = q0 + 128 < p.Nq;  // the second query tile holds rows
  const uint32_t P_COL = SPLIT ? 256 : 0, P_STRIDE = SPLIT ? 64 : 128;
  const uint32_t O_COL = SPLIT ? 384 : 256, O_STRIDE = SPLIT ? 64 : p.dn;

  if (threadIdx.x == 288) {
    for (int t = 0; t < 2; ++t) {
      mbar_init(smem_u32(&bar_q[t]), 1);
      mbar_init(smem_u32(&bar_s[t]), 1);
      mbar_init(smem_u32(&bar_p[t]), 128);
      mbar_init(smem_u32(&bar_o[t]), 1);
      mbar_init(smem_u32(&bar_sfree[t]), 128);
      mbar_init(smem_u32(&bar_pfree[t]), 1);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&kv_full[s]), 1);
      mbar_init(smem_u32(&kv_empty[s]), two ? 2 : 1);  // one commit per MMA warp
    }
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc_rt(smem_u32(tmem_slot), 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 8) {
    // ---------------------------------------------------------------- TMA producer
    if (elect_one()) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      for (int t = 0; t < (two ? 2 : 1); ++t) {
        mbar_expect_tx(smem_u32(&bar_q[t]), TILE);
        for (int x = 0; x < NB; ++x)
          tma_load_4d(smem_u32(sQ + t * TILE + x * ABOX), &tmQ, smem_u32(&bar_q[t]), x * 64, h, q0 + t * 128, b);
      }
    }
    __syncwarp();
    for (int j = 0; j < n_kv; ++j) {
      const int s = j % STAGES;
      const uint32_t ph = (j / STAGES) & 1;
      mbar_wait(smem_u32(&kv_empty[s]), ph ^ 1);
      if (elect_one()) {
        const uint32_t fb = smem_u32(&kv_full[s]);
        mbar_expect_tx(fb, 2 * TILE);
        for (int x = 0; x < NB; ++x) {
          tma_load_4d(smem_u32(sK + s * TILE + x * ABOX), &tmK, fb, x * 64, h, j * 128, b);
          tma_load_4d(smem_u32(sV + s * TILE + x * ABOX), &tmV, fb, x * 64, h, j * 128, b);
        }
      }
      __syncwarp();
    }
  } else if (warp == 9 || (warp == 10 && two)) {
    // ---------------------------------------------------------------- MMA issuers: warp 9 -> tile 0, warp 10 -> tile 1
    // (one issuing warp per query tile: a single warp serialising  wait, S_0', wait, PV_0, wait, S_1', wait, PV_1  was
    // measured busy for the whole 2900-cycle iteration -- every mbarrier wake-up and issue burst of that lone warp
    // costs 80-150 cycles more than the tensor pipe needs -- and was the bottleneck of the kernel)
    const int t = warp - 9;
    const uint32_t idesc_s = umma_idesc_f16(128, 128, 0, 0);
    const uint32_t idesc_o = umma_idesc_f16(128, p.dn, 0, 1);  // B = V is MN-major
    const int nks = p.dn / 16;
    const uint32_t hi = umma_desc_hi_sw128(1024);
    const uint32_t q_lo = umma_desc_lo(smem_u32(sQ + t * TILE), 16);
    const uint32_t k_lo0 = umma_desc_lo(smem_u32(sK), 16);       // stage 0; stage s adds s * TILE / 16
    const uint32_t v_lo0 = umma_desc_lo(smem_u32(sV), ABOX);
    const uint32_t s_tm = tmem + t * 128, p_tm = tmem + P_COL + t * P_STRIDE, o_tm = tmem + O_COL + t * O_STRIDE;
    // The issue sequences are straight-line code with compile-time operand offsets (the KV loop is unrolled over the
    // ring stages, the k-steps over their maximum with a predicate): measured with runtime loops and per-step
    // descriptor arithmetic, this lone warp needed ~480 cycles to issue 3 MMAs + a commit.
    auto issue_s = [&](uint32_t k_lo) {  // S_t = Q_t K^T
#pragma unroll
      for (int ks = 0; ks < NB * 4; ++ks) {
        const uint32_t off = ((ks >> 2) * ABOX + (ks & 3) * 32) >> 4;
        if (ks < nks) umma_f16_ss(s_tm, umma_desc_pack(q_lo + off, hi), umma_desc_pack(k_lo + off, hi), idesc_s, ks > 0);
      }
      umma_commit(smem_u32(&bar_s[t]));
    };
    mbar_wait(smem_u32(&kv_full[0]), 0);
    mbar_wait(smem_u32(&bar_q[t]), 0);
    tc_fence_after();
    if (elect_one()) issue_s(k_lo0);
    __syncwarp();
    for (int j0 = 0; j0 < n_kv; j0 += STAGES) {
#pragma unroll
      for (int s = 0; s < STAGES; ++s) {
        const int j = j0 + s;
        if (j >= n_kv) break;
        constexpr int dummy = 0;
        (void)dummy;
        const int s1 = (s + 1) % STAGES;
        const bool more = j + 1 < n_kv;
        const uint32_t v_lo = v_lo0 + s * (TILE >> 4);
        const uint32_t k_lo1 = k_lo0 + s1 * (TILE >> 4);
        const uint32_t ph = (j0 / STAGES) & 1;                       // phase of ring slot s at iteration j
        const uint32_t ph1 = s1 == 0 ? ph ^ 1 : ph;                  // ... of slot s1 at iteration j + 1
        // keys past Nk have P == 0: skip their k-steps
        const int kvalid = min(128, p.Nk - j * 128);
        if (t == 0) TB_TRACE(j, 0);
        if (more) mbar_wait(smem_u32(&kv_full[s1]), ph1);
        if (SPLIT && more) {  // the next scores of this tile, as soon as the current ones have been read
          mbar_wait(smem_u32(&bar_sfree[t]), j & 1);
          tc_fence_after();
          if (t == 0) TB_TRACE(j, 1);  // S_0(j) free
          if (elect_one()) issue_s(k_lo1);
          __syncwarp();
          if (t == 0) TB_TRACE(j, 2);  // S_0(j+1) issued
        }
        mbar_wait(smem_u32(&bar_p[t]), j & 1);
        tc_fence_after();
        if (t == 0) TB_TRACE(j, 4);  // P_0(j) seen by the MMA warp
        if (elect_one()) {
          if (kvalid == 128) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              umma_f16_ts(o_tm, p_tm + ks * 8, umma_desc_pack(v_lo + ks * 128, hi), idesc_o, (j > 0) || (ks > 0));
          } else {
            const int nkk = (kvalid + 15) >> 4;
            for (int ks = 0; ks < nkk; ++ks)
              umma_f16_ts(o_tm, p_tm + ks * 8, umma_desc_pack(v_lo + ks * 128, hi), idesc_o, (j > 0) || (ks > 0));
          }
          if (SPLIT) umma_commit(smem_u32(&bar_pfree[t]));
          if (!SPLIT && more) issue_s(k_lo1);
          if (!more) umma_commit(smem_u32(&bar_o[t]));
          umma_commit(smem_u32(&kv_empty[s]));
        }
        __syncwarp();
        if (t == 0) TB_TRACE(j, 6);  // PV_0(j) + commits issued
      }
    }
  } else if (warp < 8 && (warp < 4 || two)) {
    // ---------------------------------------------------------------- softmax warps: one thread per query row
    const int t = warp >> 2, quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    const uint32_t s_col = lane_addr + t * 128;
    const uint32_t p_col = lane_addr + P_COL + t * P_STRIDE;
    const uint32_t o_col = lane_addr + O_COL + t * O_STRIDE;
    const uint64_t sc2 = pack_f32x2(p.scale_log2, p.scale_log2);
    float m_run = 0.f, l_run = 0.f;
    if (TB_ATTN_FWD2_SKEW > 0 && t == 1) {
      const long long t_start = clock64();
      while (clock64() - t_start < TB_ATTN_FWD2_SKEW) {
      }
    }
    // one KV tile; MASKED (a compile-time flag, so the full tiles carry no per-element selects) = the last tile when
    // Nk is not a multiple of 128: columns >= limit are padding keys
    auto tile = [&](int j, auto masked_tag) {
      constexpr bool MASKED = decltype(masked_tag)::value;
      const int limit = p.Nk - j * 128;
      TB_TRACE(j, 3);
      mbar_wait(smem_u32(&bar_s[t]), j & 1);
      tc_fence_after();
      TB_TRACE(j, 4);
      // ---- pass 1: row maximum
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
      {
        uint32_t ra[32], rb[32];
        tmem_ld32(s_col, ra);
        tmem_ld_wait32(ra);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t* cur = (c & 1) ? rb : ra;
          uint32_t* nxt = (c & 1) ? ra : rb;
          if (c < 3) tmem_ld32(s_col + (c + 1) * 32, nxt);
          if (MASKED) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i >= limit) cur[i] = 0xff800000u;
          }
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            mx0 = fmax3(mx0, __uint_as_float(cur[i]), __uint_as_float(cur[i + 1]));
            mx1 = fmax3(mx1, __uint_as_float(cur[i + 2]), __uint_as_float(cur[i + 3]));
            mx2 = fmax3(mx2, __uint_as_float(cur[i + 4]), __uint_as_float(cur[i + 5]));
            mx3 = fmax3(mx3, __uint_as_float(cur[i + 6]), __uint_as_float(cur[i + 7]));
          }
          if (c < 3) tmem_ld_wait32(nxt);
        }
      }
      const float m_tile = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * p.scale_log2;
      TB_TRACE(j, 5);
      if (j == 0) {
        m_run = m_tile;
      } else {
        // lazy rescale: keep the old reference maximum unless the row maximum grew by more than 2^8
        const bool need = m_tile > m_run + 8.f;
        if (__any_sync(0xffffffffu, need)) {
          if (SPLIT) {  // PV_t(j-1) must have retired before O_t is rescaled
            mbar_wait(smem_u32(&bar_pfree[t]), (j - 1) & 1);
            tc_fence_after();
          }
          const float alpha = need ? exp2f(m_run - m_tile) : 1.f;
          // (!SPLIT: S_t(j) complete implies PV_t(j-1) complete by issue order)
          for (int c = 0; c < p.dn; c += 16) {
            uint32_t v[16];
            tmem_ld16(o_col + c, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st16(o_col + c, v);
          }
          l_run *= alpha;
        }
        if (need) m_run = m_tile;
      }
      // ---- pass 2: P = exp2(S*c - m), packed fp16
      const uint64_t negm2 = pack_f32x2(-m_run, -m_run);
      float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
      {
        uint32_t ra[32], rb[32];
        tmem_ld32(s_col, ra);
        tmem_ld_wait32(ra);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t* cur = (c & 1) ? rb : ra;
          uint32_t* nxt = (c & 1) ? ra : rb;
          if (c < 3) tmem_ld32(s_col + (c + 1) * 32, nxt);
          if (MASKED) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i >= limit) cur[i] = 0xff800000u;
          }
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint64_t e = fma_f32x2(pack_f32x2(__uint_as_float(cur[2 * i]), __uint_as_float(cur[2 * i + 1])), sc2, negm2);
            float p0, p1;
            if ((i & 7) >= 8 - TB_ATTN_FWD2_POLY) exp2_pair_poly(e, p0, p1);
            else exp2_pair_mufu(e, p0, p1);
            if (i & 1) { l2 += p0; l3 += p1; } else { l0 += p0; l1 += p1; }
            asm("cvt.rn." TB_H16X2 ".f32 %0, %1, %2;" : "=r"(pk[i]) : "f"(p1), "f"(p0));
          }
          if (c < 3) tmem_ld_wait32(nxt);
          if (SPLIT && c == 2) {  // the last chunk of S_t(j) is in registers: the next scores may overwrite it
            tc_fence_before();
            mbar_arrive(smem_u32(&bar_sfree[t]));
          }
          if (SPLIT && c == 0 && j > 0) {  // PV_t(j-1) has retired: P_t may be overwritten (waited for as late as
            mbar_wait(smem_u32(&bar_pfree[t]), (j - 1) & 1);  // possible: a quarter of the exp pass is done by now)
            tc_fence_after();
          }
          // !SPLIT: chunk c+1 (columns 32c+32 ..) is already in registers: P columns [16c, 16c+16) only cover
          // consumed scores
          tmem_st16(p_col + c * 16, pk);
        }
      }
      l_run += (l0 + l1) + (l2 + l3);
      TB_TRACE(j, 6);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(&bar_p[t]));
      TB_TRACE(j, 7);
    };
    const int n_full = p.Nk / 128;  // tiles without padding keys
    for (int j = 0; j < n_full; ++j) tile(j, std::false_type{});
    if (n_full < n_kv) tile(n_full, std::true_type{});
    // ---- epilogue: O / l
    mbar_wait(smem_u32(&bar_o[t]), 0);
    tc_fence_after();
    const float inv_l = 1.f / l_run;
    const int q = q0 + t * 128 + row;
    const bool ok = q < p.Nq;
    tb::half_t* op = p.O + ((long long)b * p.Nq + q) * p.ldo + h * p.d;
    for (int c = 0; c < p.dn; c += 16) {
      uint32_t v[16];
      tmem_ld16(o_col + c, v);
      tmem_ld_wait();
      if (ok) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (c + g * 8 < p.d) {
            uint4 o;
            o.x = pack_half2(__uint_as_float(v[g * 8 + 0]) * inv_l, __uint_as_float(v[g * 8 + 1]) * inv_l);
            o.y = pack_half2(__uint_as_float(v[g * 8 + 2]) * inv_l, __uint_as_float(v[g * 8 + 3]) * inv_l);
            o.z = pack_half2(__uint_as_float(v[g * 8 + 4]) * inv_l, __uint_as_float(v[g * 8 + 5]) * inv_l);
            o.w = pack_half2(__uint_as_float(v[g * 8 + 6]) * inv_l, __uint_as_float(v[g * 8 + 7]) * inv_l);
            *reinterpret_cast<uint4*>(op + c + g * 8) = o;
          }
        }
      }
    }
    if (ok && p.lse) p.lse[((long long)b * p.heads + h) * p.Nq + q] = m_run + log2f(l_run);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc_rt(tmem, 512u);
}

// ------------------------------------------------------------------------------------ forward, v2 with two threads per row
// Same schedule as attn_fwd2_kernel (two query tiles per CTA, P in its own TMEM columns and S_t(j+1) issued as soon
// as S_t(j) has been read when head_dim <= 64, one MMA-issuing warp per tile), but SIXTEEN softmax warps: warp w ->
// tile w >> 3, TMEM lane quarter w & 3, column half (w >> 2) & 1, i.e. two threads per row with 64 scores each.
// ncu on the one-thread-per-row kernel (profiles/r02_*): issue slots 51 % busy, XU 51 %, no pipe saturated -- two
// warps per scheduler cannot cover the fixed-latency ("wait" 29 %), MUFU-result and MIO stalls of this instruction
// stream.  Four warps per scheduler can; the price is the exchange of the two half-row maxima through shared
// memory, synchronised by a 64-thread named barrier PER WARP PAIR (not per tile), and a single pass over the scores
// (64 registers) instead of two.
__device__ __forceinline__ void st_shared_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float ld_shared_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

template <int NB, int STAGES>
__global__ void __launch_bounds__(608, 1)
attn_fwd2h_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  constexpr int TILE = NB * ABOX;
  constexpr bool SPLIT = NB == 1;  // P in its own TMEM columns
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;                  // two query tiles
  uint8_t* sK = sQ + 2 * TILE;
  uint8_t* sV = sK + STAGES * TILE;
  float* sX = reinterpret_cast<float*>(sV + STAGES * TILE);  // [2 parities][2 tiles][2 halves][128] half-row maxima / sums
  uint64_t* bars = reinterpret_cast<uint64_t*>(sX + 2 * 2 * 2 * 128);
  uint64_t* bar_q = bars;              // [2]  Q_t landed
  uint64_t* bar_s = bars + 2;          // [2]  S_t(j) complete                      (MMA -> softmax)
  uint64_t* bar_p = bars + 4;          // [2]  P_t(j) stored                        (softmax -> MMA)
  uint64_t* bar_o = bars + 6;          // [2]  last PV_t complete
  uint64_t* bar_sfree = bars + 8;      // [2]  S_t(j) is in registers               (softmax -> MMA, SPLIT)
  uint64_t* bar_pfree = bars + 10;     // [2]  PV_t(j) complete: P_t / O_t are free (MMA -> softmax, SPLIT)
  uint64_t* kv_full = bars + 12;
  uint64_t* kv_empty = bars + 12 + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12 + 2 * STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256, h = blockIdx.y, b = blockIdx.z;
  const int n_kv = p.n_inner;
  const bool two = q0 + 128 < p.Nq;  // the second query tile holds rows
  const uint32_t P_COL = SPLIT ? 256 : 0, P_STRIDE = SPLIT ? 64 : 128;
  const uint32_t O_COL = SPLIT ? 384 : 256, O_STRIDE = SPLIT ? 64 : p.dn;

  if (threadIdx.x == 512) {
    for (int t = 0; t < 2; ++t) {
      mbar_init(smem_u32(&bar_q[t]), 1);
      mbar_init(smem_u32(&bar_s[t]), 1);
      mbar_init(smem_u32(&bar_p[t]), 256);
      mbar_init(smem_u32(&bar_o[t]), 1);
      mbar_init(smem_u32(&bar_sfree[t]), 256);
      mbar_init(smem_u32(&bar_pfree[t]), 1);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&kv_full[s]), 1);
      mbar_init(smem_u32(&kv_empty[s]), two ? 2 : 1);  // one commit per MMA warp
    }
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc_rt(smem_u32(tmem_slot), 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 16) {
    // ---------------------------------------------------------------- TMA producer
    if (elect_one()) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      for (int t = 0; t < (two ? 2 : 1); ++t) {
        mbar_expect_tx(smem_u32(&bar_q[t]), TILE);
        for (int x = 0; x < NB; ++x)
          tma_load_4d(smem_u32(sQ + t * TILE + x * ABOX), &tmQ, smem_u32(&bar_q[t]), x * 64, h, q0 + t * 128, b);
      }
    }
    __syncwarp();
    for (int j = 0; j < n_kv; ++j) {
      const int s = j % STAGES;
      const uint32_t ph = (j / STAGES) & 1;
      mbar_wait(smem_u32(&kv_empty[s]), ph ^ 1);
      if (elect_one()) {
        const uint32_t fb = smem_u32(&kv_full[s]);
        mbar_expect_tx(fb, 2 * TILE);
        for (int x = 0; x < NB; ++x) {
          tma_load_4d(smem_u32(sK + s * TILE + x * ABOX), &tmK, fb, x * 64, h, j * 128, b);
          tma_load_4d(smem_u32(sV + s * TILE + x * ABOX), &tmV, fb, x * 64, h, j * 128, b);
        }
      }
      __syncwarp();
    }
  } else if (warp == 17 || (warp == 18 && two)) {
    // ---------------------------------------------------------------- MMA issuers: warp 17 -> tile 0, warp 18 -> tile 1
    const int t = warp - 17;
    const uint32_t idesc_s = umma_idesc_f16(128, 128, 0, 0);
    const uint32_t idesc_o = umma_idesc_f16(128, p.dn, 0, 1);  // B = V is MN-major
    const int nks = p.dn / 16;
    const uint32_t hi = umma_desc_hi_sw128(1024);
    const uint32_t q_lo = umma_desc_lo(smem_u32(sQ + t * TILE), 16);
    const uint32_t k_lo0 = umma_desc_lo(smem_u32(sK), 16);       // stage 0; stage s adds s * TILE / 16
    const uint32_t v_lo0 = umma_desc_lo(smem_u32(sV), ABOX);
    const uint32_t s_tm = tmem + t * 128, p_tm = tmem + P_COL + t * P_STRIDE, o_tm = tmem + O_COL + t * O_STRIDE;
    auto issue_s = [&](uint32_t k_lo) {  // S_t = Q_t K^T
#pragma unroll
      for (int ks = 0; ks < NB * 4; ++ks) {
        const uint32_t off = ((ks >> 2) * ABOX + (ks & 3) * 32) >> 4;
        if (ks < nks) umma_f16_ss(s_tm, umma_desc_pack(q_lo + off, hi), umma_desc_pack(k_lo + off, hi), idesc_s, ks > 0);
      }
      umma_commit(smem_u32(&bar_s[t]));
    };
    mbar_wait(smem_u32(&kv_full[0]), 0);
    mbar_wait(smem_u32(&bar_q[t]), 0);
    tc_fence_after();
    if (elect_one()) issue_s(k_lo0);
    __syncwarp();
    for (int j0 = 0; j0 < n_kv; j0 += STAGES) {
#pragma unroll
      for (int s = 0; s < STAGES; ++s) {
        const int j = j0 + s;
        if (j >= n_kv) break;
        const int s1 = (s + 1) % STAGES;
        const bool more = j + 1 < n_kv;
        const uint32_t v_lo = v_lo0 + s * (TILE >> 4);
        const uint32_t k_lo1 = k_lo0 + s1 * (TILE >> 4);
        const uint32_t ph = (j0 / STAGES) & 1;                       // phase of ring slot s at iteration j
        const uint32_t ph1 = s1 == 0 ? ph ^ 1 : ph;                  // ... of slot s1 at iteration j + 1
        const int kvalid = min(128, p.Nk - j * 128);                 // keys past Nk have P == 0: skip their k-steps
        if (more) mbar_wait(smem_u32(&kv_full[s1]), ph1);
        if (SPLIT && more) {  // the next scores of this tile, as soon as the current ones have been read
          mbar_wait(smem_u32(&bar_sfree[t]), j & 1);
          tc_fence_after();
          if (elect_one()) issue_s(k_lo1);
          __syncwarp();
        }
        mbar_wait(smem_u32(&bar_p[t]), j & 1);
        tc_fence_after();
        if (elect_one()) {
          if (kvalid == 128) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              umma_f16_ts(o_tm, p_tm + ks * 8, umma_desc_pack(v_lo + ks * 128, hi), idesc_o, (j > 0) || (ks > 0));
          } else {
            const int nkk = (kvalid + 15) >> 4;
            for (int ks = 0; ks < nkk; ++ks)
              umma_f16_ts(o_tm, p_tm + ks * 8, umma_desc_pack(v_lo + ks * 128, hi), idesc_o, (j > 0) || (ks > 0));
          }
          if (SPLIT) umma_commit(smem_u32(&bar_pfree[t]));
          if (!SPLIT && more) issue_s(k_lo1);
          if (!more) umma_commit(smem_u32(&bar_o[t]));
          umma_commit(smem_u32(&kv_empty[s]));
        }
        __syncwarp();
      }
    }
  } else if (warp < 16 && (warp < 8 || two)) {
    // ---------------------------------------------------------------- softmax warps: two threads per query row
    const int t = warp >> 3, quarter = warp & 3, half = (warp >> 2) & 1;
    const int row = quarter * 32 + lane;
    const int pair_bar = 1 + t * 4 + quarter;  // named barrier of the two warps that share these 32 rows
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    const uint32_t s_col = lane_addr + t * 128 + half * 64;
    const uint32_t p_col = lane_addr + P_COL + t * P_STRIDE + half * 32;
    const uint32_t o_col = lane_addr + O_COL + t * O_STRIDE;
    const uint64_t sc2 = pack_f32x2(p.scale_log2, p.scale_log2);
    float m_run = 0.f, l_run = 0.f;  // l_run: this thread's 64 columns only
    const uint32_t sx_base = smem_u32(sX);
    auto tile = [&](int j, auto masked_tag) {
      constexpr bool MASKED = decltype(masked_tag)::value;
      const int limit = p.Nk - j * 128 - half * 64;  // this thread's columns >= limit are padding keys
      const uint32_t xmine = sx_base + ((((j & 1) * 2 + t) * 2 + half) * 128 + row) * 4;
      const uint32_t xother = sx_base + ((((j & 1) * 2 + t) * 2 + (half ^ 1)) * 128 + row) * 4;
      mbar_wait(smem_u32(&bar_s[t]), j & 1);
      tc_fence_after();
      uint32_t r[64];
      tmem_ld32(s_col, r);
      tmem_ld32(s_col + 32, r + 32);
      tmem_ld_wait();
      if (SPLIT) {  // S_t(j) is in registers: the next scores may overwrite it
        tc_fence_before();
        mbar_arrive(smem_u32(&bar_sfree[t]));
      }
      if (MASKED) {
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i >= limit) r[i] = 0xff800000u;
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 64; i += 8) {
        mx0 = fmax3(mx0, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
        mx1 = fmax3(mx1, __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
        mx2 = fmax3(mx2, __uint_as_float(r[i + 4]), __uint_as_float(r[i + 5]));
        mx3 = fmax3(mx3, __uint_as_float(r[i + 6]), __uint_as_float(r[i + 7]));
      }
      float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      st_shared_f32(xmine, mx);
      named_bar_sync(pair_bar, 64);  // (!SPLIT: both halves of these rows hold their scores: P may overwrite S)
      mx = fmaxf(mx, ld_shared_f32(xother));
      const float m_tile = mx * p.scale_log2;
      if (j == 0) {
        m_run = m_tile;
      } else {
        // lazy rescale: keep the old reference maximum unless the row maximum grew by more than 2^8
        const bool need = m_tile > m_run + 8.f;
        if (__any_sync(0xffffffffu, need)) {
          if (SPLIT) {  // PV_t(j-1) must have retired before O_t is rescaled
            mbar_wait(smem_u32(&bar_pfree[t]), (j - 1) & 1);
            tc_fence_after();
          }
          const float alpha = need ? exp2f(m_run - m_tile) : 1.f;
          // (!SPLIT: S_t(j) complete implies PV_t(j-1) complete by issue order.)  The halves alternate 16-column chunks.
          for (int c = half * 16; c < p.dn; c += 32) {
            uint32_t v[16];
            tmem_ld16(o_col + c, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st16(o_col + c, v);
          }
          l_run *= alpha;
        }
        if (need) m_run = m_tile;
      }
      // ---- P = exp2(S*c - m), packed fp16
      const uint64_t negm2 = pack_f32x2(-m_run, -m_run);
      uint64_t la = pack_f32x2(0.f, 0.f), lb = pack_f32x2(0.f, 0.f);
      uint32_t pk[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const uint64_t e = fma_f32x2(pack_f32x2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), sc2, negm2);
        float p0, p1;
        if ((i & 7) >= 8 - TB_ATTN_FWD2_POLY) exp2_pair_poly(e, p0, p1);
        else exp2_pair_mufu(e, p0, p1);
        if (i & 1) lb = add_f32x2(lb, pack_f32x2(p0, p1));
        else la = add_f32x2(la, pack_f32x2(p0, p1));
        asm("cvt.rn." TB_H16X2 ".f32 %0, %1, %2;" : "=r"(pk[i]) : "f"(p1), "f"(p0));
      }
      {
        float a0, a1, b0, b1;
        unpack_f32x2(la, a0, a1);
        unpack_f32x2(lb, b0, b1);
        l_run += (a0 + a1) + (b0 + b1);
      }
      if (SPLIT && j > 0) {  // PV_t(j-1) has retired: P_t may be overwritten
        mbar_wait(smem_u32(&bar_pfree[t]), (j - 1) & 1);
        tc_fence_after();
      }
      tmem_st32(p_col, pk);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(&bar_p[t]));
    };
    const int n_full = p.Nk / 128;  // tiles without padding keys
    for (int j = 0; j < n_full; ++j) tile(j, std::false_type{});
    if (n_full < n_kv) tile(n_full, std::true_type{});
    // ---- epilogue: O / l
    {
      const uint32_t xmine = sx_base + ((((n_kv & 1) * 2 + t) * 2 + half) * 128 + row) * 4;
      const uint32_t xother = sx_base + ((((n_kv & 1) * 2 + t) * 2 + (half ^ 1)) * 128 + row) * 4;
      st_shared_f32(xmine, l_run);
      named_bar_sync(pair_bar, 64);
      l_run += ld_shared_f32(xother);
    }
    mbar_wait(smem_u32(&bar_o[t]), 0);
    tc_fence_after();
    const float inv_l = 1.f / l_run;
    const int q = q0 + t * 128 + row;
    const bool ok = q < p.Nq;
    tb::half_t* op = p.O + ((long long)b * p.Nq + q) * p.ldo + h * p.d;
    for (int c = half * 16; c < p.dn; c += 32) {
      uint32_t v[16];
      tmem_ld16(o_col + c, v);
      tmem_ld_wait();
      if (ok) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (c + g * 8 < p.d) {
            uint4 o;
            o.x = pack_half2(__uint_as_float(v[g * 8 + 0]) * inv_l, __uint_as_float(v[g * 8 + 1]) * inv_l);
            o.y = pack_half2(__uint_as_float(v[g * 8 + 2]) * inv_l, __uint_as_float(v[g * 8 + 3]) * inv_l);
            o.z = pack_half2(__uint_as_float(v[g * 8 + 4]) * inv_l, __uint_as_float(v[g * 8 + 5]) * inv_l);
            o.w = pack_half2(__uint_as_float(v[g * 8 + 6]) * inv_l, __uint_as_float(v[g * 8 + 7]) * inv_l);
            *reinterpret_cast<uint4*>(op + c + g * 8) = o;
          }
        }
      }
    }
    if (ok && half == 0 && p.lse) p.lse[((long long)b * p.heads + h) * p.Nq + q] = m_run + log2f(l_run);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc_rt(tmem, 512u);
}

// ------------------------------------------------------------------------------------ backward
// delta[b,h,q] = sum_c dO[b,q,h*d+c] * O[b,q,h*d+c]   (one warp per (b,q), lanes over heads*d/8 vectors)
__global__ void attn_delta_kernel(const tb::half_t* __restrict__ O, const tb::half_t* __restrict__ dO,
                                  long long ldo, long long lddo, float* __restrict__ delta, int B,
                                  int Nq, int heads, int d) {
  // all 32 lanes stream the row (one 16-byte vector of O and dO each per trip); the per-vector partial dot
  // products meet in shared memory and lane h sums head h's d/8 vectors.  (Looping over heads with only d/8
  // lanes active, as before, left 27 of 32 lanes idle at d = 40.)
  __shared__ float part[8][160];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + w;
  if (row >= (long long)B * Nq) return;
  const int b = row / Nq, q = row % Nq;
  const int vph = d / 8, nvec = heads * vph;
  for (int v = lane; v < nvec; v += 32) {
    const uint4 a = *reinterpret_cast<const uint4*>(O + row * ldo + v * 8);
    const uint4 g = *reinterpret_cast<const uint4*>(dO + row * lddo + v * 8);
    const tb::half2_t* ah = reinterpret_cast<const tb::half2_t*>(&a);
    const tb::half2_t* gh = reinterpret_cast<const tb::half2_t*>(&g);
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 x = tb::h22f2(ah[i]), y = tb::h22f2(gh[i]);
      acc += x.x * y.x + x.y * y.y;
    }
    part[w][v] = acc;
  }
  __syncwarp();
  for (int hh = lane; hh < heads; hh += 32) {
    float acc = 0.f;
    for (int v = 0; v < vph; ++v) acc += part[w][hh * vph + v];
    delta[((long long)b * heads + hh) * Nq + q] = acc;
  }
}

// q-split backward: fp32 accumulators [2][rows][C] (dV, dK) -> fp16 dV / dK with their row strides
__global__ void attn_dkv_finish_kernel(const float* __restrict__ ws, tb::half_t* __restrict__ dK, long long lddk,
                                       tb::half_t* __restrict__ dV, long long lddv, long long rows, int C) {
  const long long nvec = rows * (C / 4);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * nvec;
       i += (long long)gridDim.x * blockDim.x) {
    const int which = i >= nvec;
    const long long j = which ? i - nvec : i;
    const long long r = j / (C / 4);
    const int c = (int)(j % (C / 4)) * 4;
    const float4 f = *reinterpret_cast<const float4*>(ws + (which * rows + r) * C + c);
    tb::half_t* dst = which ? dK + r * lddk + c : dV + r * lddv + c;
    uint2 o;
    o.x = pack_half2(f.x, f.y);
    o.y = pack_half2(f.z, f.w);
    *reinterpret_cast<uint2*>(dst) = o;
  }
}

// CTA owns one 128-row KV tile of one (b, h) and loops over the Q tiles.
//   S^T = K Q^T ; P^T = exp2(S^T*c - L) ; dP^T = V dO^T ; dS^T = P^T o (dP^T - delta)
//   dV += P^T dO ; dK += dS^T Q ; dQ_i = dS K  (red.add into the fp32 dQ accumulator)
// TMEM: X = [0,128) is S^T, then dP^T, then dQ_i in turn; dV at 128, dK at 128+dn.
// Shared memory: dS^T overwrites P^T in place (P^T stays in registers for the dS product), which keeps
// the head_dim <= 64 kernel at 98 KB so two CTAs co-reside per SM.
template <int NB, int STAGES>
__global__ void __launch_bounds__(320, NB == 1 ? 2 : 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                const AttnParams p) {
  constexpr int TILE = NB * ABOX;
  constexpr uint32_t X_COL = 0, DV_COL = 128;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + TILE;
  uint8_t* sQ = sV + TILE;
  uint8_t* sdO = sQ + STAGES * TILE;
  uint8_t* sPT = sdO + STAGES * TILE;  // P^T, then dS^T in place
  uint8_t* sDS = sPT;
  float* sL = reinterpret_cast<float*>(sPT + 2 * ABOX);  // [128] L_i, [128] delta_i
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
  // blockIdx.x = kv tile * qsplit + q split: under-filled launches (cross attention: one KV tile per (b,h)) are
  // split along the Q loop so the grid covers the machine; the splits meet in fp32 red.adds on dkv_ws
  const int kv0 = (blockIdx.x / p.qsplit) * 128, h = blockIdx.y, b = blockIdx.z;
  const int qs = blockIdx.x % p.qsplit;
  const uint32_t DK_COL = DV_COL + p.dn;
  const int q_per = (p.n_inner + p.qsplit - 1) / p.qsplit;
  // causal: q tiles entirely before this kv tile see none of its keys (never combined with qsplit > 1)
  const int i_begin = p.causal ? kv0 / 128 : qs * q_per;
  const int i_end = p.causal ? p.n_inner : min(p.n_inner, i_begin + q_per);

  if (threadIdx.x == 288) {
    mbar_init(smem_u32(bar_kv), 1);
    mbar_init(smem_u32(bar_s), 1);
    mbar_init(smem_u32(bar_pt), 256);
    mbar_init(smem_u32(bar_dp), 1);
    mbar_init(smem_u32(bar_dv), 1);
    mbar_init(smem_u32(bar_ds), 256);
    mbar_init(smem_u32(bar_dq), 1);
    mbar_init(smem_u32(bar_dqfree), 256);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&q_full[s]), 1);
      mbar_init(smem_u32(&q_empty[s]), 1);
    }
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc_rt(smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 8) {
    // ---------------------------------------------------------------- TMA producer (warp-uniform)
    if (elect_one()) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      tma_prefetch_desc(&tmdO);
      mbar_expect_tx(smem_u32(bar_kv), 2 * TILE);
      for (int x = 0; x < NB; ++x) {
        tma_load_4d(smem_u32(sK + x * ABOX), &tmK, smem_u32(bar_kv), x * 64, h, kv0, b);
        tma_load_4d(smem_u32(sV + x * ABOX), &tmV, smem_u32(bar_kv), x * 64, h, kv0, b);
      }
    }
    __syncwarp();
    for (int i = i_begin; i < i_end; ++i) {
      const int it = i - i_begin;
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      mbar_wait(smem_u32(&q_empty[s]), ph ^ 1);
      if (elect_one()) {
        const uint32_t fb = smem_u32(&q_full[s]);
        mbar_expect_tx(fb, 2 * TILE);
        for (int x = 0; x < NB; ++x) {
          tma_load_4d(smem_u32(sQ + s * TILE + x * ABOX), &tmQ, fb, x * 64, h, i * 128, b);
          tma_load_4d(smem_u32(sdO + s * TILE + x * ABOX), &tmdO, fb, x * 64, h, i * 128, b);
        }
      }
      __syncwarp();
    }
  } else if (warp == 9) {
    // ---------------------------------------------------------------- MMA issuer (warp-uniform)
    const uint32_t idesc_kk = umma_idesc_f16(128, 128, 0, 0);     // S^T, dP^T : both K-major
    const uint32_t idesc_acc = umma_idesc_f16(128, p.dn, 0, 1);   // dV, dK : B MN-major
    const int nks = p.dn / 16;
    // dQ = dS K: A = dS^T buffer read MN-major, B = K MN-major; N split at 128 when dn > 128
    const int dq_n0 = p.dn > 128 ? 128 : p.dn;
    const int dq_n1 = p.dn - dq_n0;
    const uint32_t idesc_dq0 = umma_idesc_f16(128, dq_n0, 1, 1);
    const uint32_t idesc_dq1 = umma_idesc_f16(128, dq_n1 > 0 ? dq_n1 : 16, 1, 1);
    const uint32_t hi = umma_desc_hi_sw128(1024);
    // K-major views (LBO unused) and MN-major views (LBO = distance between 64-column boxes)
    const uint32_t k_lo = umma_desc_lo(smem_u32(sK), 16), v_lo = umma_desc_lo(smem_u32(sV), 16);
    const uint32_t pt_lo = umma_desc_lo(smem_u32(sPT), 16);
    const uint32_t k_mn = umma_desc_lo(smem_u32(sK), ABOX), ds_mn = umma_desc_lo(smem_u32(sDS), ABOX);
    uint32_t ph_dqfree = 0;
    mbar_wait(smem_u32(bar_kv), 0);
    for (int i = i_begin; i < i_end; ++i) {
      const int it = i - i_begin;
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      const uint32_t q_lo = umma_desc_lo(smem_u32(sQ + s * TILE), 16), do_lo = umma_desc_lo(smem_u32(sdO + s * TILE), 16);
      const uint32_t q_mn = umma_desc_lo(smem_u32(sQ + s * TILE), ABOX), do_mn = umma_desc_lo(smem_u32(sdO + s * TILE), ABOX);
      mbar_wait(smem_u32(&q_full[s]), ph);
      if (it > 0) {
        mbar_wait(smem_u32(bar_dqfree), ph_dqfree);
        ph_dqfree ^= 1;
      }
      tc_fence_after();
      if (elect_one()) {
        for (int ks = 0; ks < nks; ++ks) {  // S^T = K Q^T
          const uint32_t off = ((ks >> 2) * ABOX + (ks & 3) * 32) >> 4;
          umma_f16_ss(tmem + X_COL, umma_desc_pack(k_lo + off, hi), umma_desc_pack(q_lo + off, hi), idesc_kk, ks > 0);
        }
        umma_commit(smem_u32(bar_s));
      }
      __syncwarp();
      mbar_wait(smem_u32(bar_pt), it & 1);  // P^T in smem, S^T drained from TMEM
      tc_fence_after();
      if (elect_one()) {
        for (int ks = 0; ks < nks; ++ks) {  // dP^T = V dO^T
          const uint32_t off = ((ks >> 2) * ABOX + (ks & 3) * 32) >> 4;
          umma_f16_ss(tmem + X_COL, umma_desc_pack(v_lo + off, hi), umma_desc_pack(do_lo + off, hi), idesc_kk, ks > 0);
        }
        umma_commit(smem_u32(bar_dp));
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {  // dV += P^T dO
          const uint32_t offp = ((ks >> 2) * ABOX + (ks & 3) * 32) >> 4;
          umma_f16_ss(tmem + DV_COL, umma_desc_pack(pt_lo + offp, hi), umma_desc_pack(do_mn + ks * 128, hi),
                      idesc_acc, (it > 0) || (ks > 0));
        }
        umma_commit(smem_u32(bar_dv));  // P^T buffer may now be overwritten by dS^T
      }
      __syncwarp();
      mbar_wait(smem_u32(bar_ds), it & 1);  // dS^T in smem, dP^T drained
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {  // dK += dS^T Q
          const uint32_t offp = ((ks >> 2) * ABOX + (ks & 3) * 32) >> 4;
          umma_f16_ss(tmem + DK_COL, umma_desc_pack(pt_lo + offp, hi), umma_desc_pack(q_mn + ks * 128, hi),
                      idesc_acc, (it > 0) || (ks > 0));
        }
        if (p.dQacc || p.dQ16) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {  // dQ_i = dS K   (A: [kv][q] read MN-major)
            umma_f16_ss(tmem + X_COL, umma_desc_pack(ds_mn + ks * 128, hi), umma_desc_pack(k_mn + ks * 128, hi),
                        idesc_dq0, ks > 0);
          }
        }
        umma_commit(smem_u32(bar_dq));
      }
      __syncwarp();
      if ((p.dQacc || p.dQ16) && dq_n1 > 0) {
        // columns 128..dn of dQ reuse X once the first 128 have been drained
        mbar_wait(smem_u32(bar_dqfree), ph_dqfree);
        ph_dqfree ^= 1;
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            umma_f16_ss(tmem + X_COL, umma_desc_pack(ds_mn + ks * 128, hi),
                        umma_desc_pack(k_mn + ((2 * ABOX) >> 4) + ks * 128, hi), idesc_dq1, ks > 0);
          }
          umma_commit(smem_u32(bar_dq));
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(smem_u32(&q_empty[s]));
      __syncwarp();
    }
  } else {
    // ---------------------------------------------------------------- softmax / dS / epilogue
    // two threads per kv row: warp w owns lane quarter w & 3 and the q-column half w >> 2
    const int quarter = warp & 3, half = warp >> 2;
    const int row = quarter * 32 + lane;  // kv row inside the tile (and q row for the dQ drain)
    const int cb = half * 64;             // first q column of this thread
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    const bool kv_ok = kv0 + row < p.Nk;
    const long long bh = (long long)b * p.heads + h;
    uint32_t ph_dq = 0;
    // L_i (half 0) / delta_i (half 1) of this thread's query row: loaded one q tile ahead, so the global-load
    // latency is off the per-iteration critical path (it used to sit between two barriers)
    auto load_ld = [&](int q) -> float {
      // +inf makes exp2(. - L) == 0 for padding queries
      if (half == 0) return q < p.Nq ? p.lse[bh * p.Nq + q] : INFINITY;
      return q < p.Nq ? p.delta[bh * p.Nq + q] : 0.f;
    };
    float ld_next = load_ld(i_begin * 128 + row);
    for (int i = i_begin; i < i_end; ++i) {
      const int it = i - i_begin;
      const int q0 = i * 128;
      if (half == 0) sL[row] = ld_next;
      else sD[row] = ld_next;
      if (i + 1 < i_end) ld_next = load_ld(q0 + 128 + row);
      named_bar_sync(1, 256);
      // causal: query column c sees this key row iff q0 + c >= kv0 + row
      const int cmin = p.causal ? (kv0 + row - q0) : 0;
      const bool masked = __any_sync(0xffffffffu, cmin > cb);
      mbar_wait(smem_u32(bar_s), it & 1);
      tc_fence_after();
      uint32_t pt[32];  // this thread's 64 entries of P^T as packed fp16, kept for the dS product
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld32(lane_addr + X_COL + cb + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float4 l0 = *reinterpret_cast<const float4*>(sL + cb + c * 32 + g * 8);
          const float4 l1 = *reinterpret_cast<const float4*>(sL + cb + c * 32 + g * 8 + 4);
          uint32_t* o = pt + c * 16 + g * 4;
          o[0] = exp2_pack(fmaf(__uint_as_float(r[g * 8 + 0]), p.scale_log2, -l0.x),
                           fmaf(__uint_as_float(r[g * 8 + 1]), p.scale_log2, -l0.y));
          o[1] = exp2_pack(fmaf(__uint_as_float(r[g * 8 + 2]), p.scale_log2, -l0.z),
                           fmaf(__uint_as_float(r[g * 8 + 3]), p.scale_log2, -l0.w));
          o[2] = exp2_pack(fmaf(__uint_as_float(r[g * 8 + 4]), p.scale_log2, -l1.x),
                           fmaf(__uint_as_float(r[g * 8 + 5]), p.scale_log2, -l1.y));
          o[3] = exp2_pack(fmaf(__uint_as_float(r[g * 8 + 6]), p.scale_log2, -l1.z),
                           fmaf(__uint_as_float(r[g * 8 + 7]), p.scale_log2, -l1.w));
          if (masked) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int col = cb + c * 32 + g * 8 + 2 * k;
              if (col < cmin) o[k] &= 0xffff0000u;
              if (col + 1 < cmin) o[k] &= 0x0000ffffu;
            }
          }
          if (!kv_ok) o[0] = o[1] = o[2] = o[3] = 0u;
          *reinterpret_cast<uint4*>(sPT + sw128_off(row, half * 8 + c * 4 + g)) =
              make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
      fence_async_smem();
      tc_fence_before();
      mbar_arrive(smem_u32(bar_pt));

      mbar_wait(smem_u32(bar_dp), it & 1);
      mbar_wait(smem_u32(bar_dv), it & 1);  // the dV MMA has finished reading P^T: dS^T overwrites it
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld32(lane_addr + X_COL + cb + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float4 d0 = *reinterpret_cast<const float4*>(sD + cb + c * 32 + g * 8);
          const float4 d1 = *reinterpret_cast<const float4*>(sD + cb + c * 32 + g * 8 + 4);
          const uint32_t* pp = pt + c * 16 + g * 4;
          uint4 o;
          o.x = hmul2_u32(pp[0], pack_half2(__uint_as_float(r[g * 8 + 0]) - d0.x, __uint_as_float(r[g * 8 + 1]) - d0.y));
          o.y = hmul2_u32(pp[1], pack_half2(__uint_as_float(r[g * 8 + 2]) - d0.z, __uint_as_float(r[g * 8 + 3]) - d0.w));
          o.z = hmul2_u32(pp[2], pack_half2(__uint_as_float(r[g * 8 + 4]) - d1.x, __uint_as_float(r[g * 8 + 5]) - d1.y));
          o.w = hmul2_u32(pp[3], pack_half2(__uint_as_float(r[g * 8 + 6]) - d1.z, __uint_as_float(r[g * 8 + 7]) - d1.w));
          *reinterpret_cast<uint4*>(sDS + sw128_off(row, half * 8 + c * 4 + g)) = o;
        }
      }
      fence_async_smem();
      tc_fence_before();
      mbar_arrive(smem_u32(bar_ds));

      const int n_chunks = ((p.dQacc || p.dQ16) && p.dn > 128) ? 2 : 1;
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
          // Lane pairs trade halves so that BOTH lanes of a pair hit the same 32-byte sector in one red
          // instruction (even lane: columns +0..3, odd lane: +4..7 of the same row): the L2 sees half as many
          // sector requests as with one row per lane.  The reds cost ~25 % of the kernel (measured by skipping them).
          const bool even = (lane & 1) == 0;
          const bool q_ok_a = (q0 + (row & ~1)) < p.Nq, q_ok_b = (q0 + (row | 1)) < p.Nq;
          float* dq_a = dq - (even ? 0 : p.lddq) + (even ? 0 : 4);  // row R = even row of the pair
          float* dq_b = dq + (even ? p.lddq : 0) + (even ? 0 : 4);  // row R + 1
          for (int c = half * 16; c < ncols; c += 32) {  // the two halves alternate 16-column chunks
            uint32_t r[16];
            tmem_ld16(lane_addr + X_COL + c, r);
            tmem_ld_wait();
#pragma unroll
            for (int gp = 0; gp < 2; ++gp) {
              float lo[4], hi[4], rc[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                lo[k] = __uint_as_float(r[gp * 8 + k]) * p.scale;
                hi[k] = __uint_as_float(r[gp * 8 + 4 + k]) * p.scale;
                rc[k] = __shfl_xor_sync(0xffffffffu, even ? hi[k] : lo[k], 1);
              }
              if (cbase + c + gp * 8 < p.d && !p.experiment) {
                if (q_ok_a)
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dq_a + c + gp * 8),
                               "f"(even ? lo[0] : rc[0]), "f"(even ? lo[1] : rc[1]), "f"(even ? lo[2] : rc[2]),
                               "f"(even ? lo[3] : rc[3])
                               : "memory");
                if (q_ok_b)
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dq_b + c + gp * 8),
                               "f"(even ? rc[0] : hi[0]), "f"(even ? rc[1] : hi[1]), "f"(even ? rc[2] : hi[2]),
                               "f"(even ? rc[3] : hi[3])
                               : "memory");
              }
            }
          }
        } else if (p.dQ16) {
          // single KV tile: this CTA holds the whole dQ row -> one fp16 store, no accumulator, no memset, no cast
          const int q = q0 + row;
          const int cbase = ch * 128;
          tb::half_t* dq = p.dQ16 + ((long long)b * p.Nq + q) * p.lddq16 + h * p.d + cbase;
          const int ncols = (p.dn - cbase) > 128 ? 128 : (p.dn - cbase);
          for (int c = half * 16; c < ncols; c += 32) {
            uint32_t r[16];
            tmem_ld16(lane_addr + X_COL + c, r);
            tmem_ld_wait();
            if (q < p.Nq) {
#pragma unroll
              for (int gp = 0; gp < 2; ++gp) {
                if (cbase + c + gp * 8 < p.d) {
                  uint4 o;
                  o.x = pack_half2(__uint_as_float(r[gp * 8 + 0]) * p.scale, __uint_as_float(r[gp * 8 + 1]) * p.scale);
                  o.y = pack_half2(__uint_as_float(r[gp * 8 + 2]) * p.scale, __uint_as_float(r[gp * 8 + 3]) * p.scale);
                  o.z = pack_half2(__uint_as_float(r[gp * 8 + 4]) * p.scale, __uint_as_float(r[gp * 8 + 5]) * p.scale);
                  o.w = pack_half2(__uint_as_float(r[gp * 8 + 6]) * p.scale, __uint_as_float(r[gp * 8 + 7]) * p.scale);
                  *reinterpret_cast<uint4*>(dq + c + gp * 8) = o;
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
    for (int c = half * 16; c < p.dn; c += 32) {
      uint32_t rv[16], rk[16];
      tmem_ld16(lane_addr + DV_COL + c, rv);
      tmem_ld16(lane_addr + DK_COL + c, rk);
      tmem_ld_wait();
      if (kv_ok && p.qsplit > 1) {
        const long long Cc = (long long)p.heads * p.d;
        float* wv = p.dkv_ws + ((long long)b * p.Nk + kv0 + row) * Cc + h * p.d;
        float* wk = wv + (long long)gridDim.z * p.Nk * Cc;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (c + g * 4 < p.d) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(wv + c + g * 4),
                         "f"(__uint_as_float(rv[g * 4 + 0])), "f"(__uint_as_float(rv[g * 4 + 1])),
                         "f"(__uint_as_float(rv[g * 4 + 2])), "f"(__uint_as_float(rv[g * 4 + 3]))
                         : "memory");
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(wk + c + g * 4),
                         "f"(__uint_as_float(rk[g * 4 + 0]) * p.scale), "f"(__uint_as_float(rk[g * 4 + 1]) * p.scale),
                         "f"(__uint_as_float(rk[g * 4 + 2]) * p.scale), "f"(__uint_as_float(rk[g * 4 + 3]) * p.scale)
                         : "memory");
          }
        }
      } else if (kv_ok) {
        tb::half_t* dv = p.dV + ((long long)b * p.Nk + kv0 + row) * p.lddv + h * p.d;
        tb::half_t* dk = p.dK + ((long long)b * p.Nk + kv0 + row) * p.lddk + h * p.d;
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
  if (warp == 0) tmem_dealloc_rt(tmem, (uint32_t)p.tmem_cols);
}


// ------------------------------------------------------------------------------------ backward, v2
// head_dim <= 64, non-causal (every UNet attention).  ONE CTA per SM with 512 TMEM columns and 576 threads:
// warps 0..15 = softmax / dS / drain (four threads per KV row, 32 query columns each), warp 16 = TMA, warp 17 =
// MMA.  attn_bwd_kernel runs its five MMAs and three softmax phases of a (kv, q) pair strictly one after the
// other through a single TMEM buffer and relies on a second resident CTA for overlap (measured: 5.7k cycles per
// pair per SM against ~1k of MUFU and ~1k of tensor work).  Here nothing the softmax warps need is ever produced
// on demand:
//   S^T(i+1) is issued as soon as the P phase has READ S^T(i)          (own buffer, [0,128))
//   dP^T(i+1) is issued as soon as the dS phase has READ dP^T(i)       (own buffer, [128,256))
//   dQ(i) has its own accumulator [256,320) and is drained after the P phase of pair i+1
//   P^T and dS^T live in separate shared-memory tiles, so dV(i) and dK/dQ(i) never block the next phase
// so the warps run P(i+1) -> drain dQ(i) -> dS(i+1) back to back.  dV at [320,384), dK at [384,448).
template <int STAGES>
__global__ void __launch_bounds__(576, 1)
attn_bwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                 const __grid_constant__ CUtensorMap tmDQ, const AttnParams p) {
  constexpr int TILE = ABOX;
  constexpr uint32_t S_COL = 0, DP_COL = 128, DQ_COL = 256, DV_COL = 320, DK_COL = 384;
  constexpr int NSM = 512;  // softmax threads

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + TILE;
  uint8_t* sQ = sV + TILE;
  uint8_t* sdO = sQ + STAGES * TILE;
  uint8_t* sPT = sdO + STAGES * TILE;  // P^T  [128 kv x 128 q] fp16, two swizzled boxes
  uint8_t* sDS = sPT + 2 * ABOX;       // dS^T, same layout
  float* sDQ = reinterpret_cast<float*>(sDS + 2 * ABOX);  // [128 q][d] fp32 staging tile of the dQ TMA reduce-add
  float* sL = sDQ + 128 * 64;                             // [2][128] L_i   (double-buffered by pair parity)
  float* sD = sL + 256;                                  // [2][128] delta_i
  uint64_t* bars = reinterpret_cast<uint64_t*>(sD + 256);
  uint64_t* bar_kv = bars;
  uint64_t* bar_s = bars + 1;
  uint64_t* bar_sfree = bars + 2;
  uint64_t* bar_pt = bars + 3;
  uint64_t* bar_dv = bars + 4;
  uint64_t* bar_dp = bars + 5;
  uint64_t* bar_dpfree = bars + 6;
  uint64_t* bar_ds = bars + 7;
  uint64_t* bar_dq = bars + 8;
  uint64_t* bar_dqfree = bars + 9;
  uint64_t* q_full = bars + 10;
  uint64_t* q_empty = bars + 10 + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10 + 2 * STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kv0 = (blockIdx.x / p.qsplit) * 128, h = blockIdx.y, b = blockIdx.z;
  const int qs = blockIdx.x % p.qsplit;
  const int q_per = (p.n_inner + p.qsplit - 1) / p.qsplit;
  const int i_begin = qs * q_per;
  const int i_end = min(p.n_inner, i_begin + q_per);

  if (threadIdx.x == NSM) {
    mbar_init(smem_u32(bar_kv), 1);
    mbar_init(smem_u32(bar_s), 1);
    mbar_init(smem_u32(bar_sfree), NSM);
    mbar_init(smem_u32(bar_pt), NSM);
    mbar_init(smem_u32(bar_dv), 1);
    mbar_init(smem_u32(bar_dp), 1);
    mbar_init(smem_u32(bar_dpfree), NSM);
    mbar_init(smem_u32(bar_ds), NSM);
    mbar_init(smem_u32(bar_dq), 1);
    mbar_init(smem_u32(bar_dqfree), NSM);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&q_full[s]), 1);
      mbar_init(smem_u32(&q_empty[s]), 1);
    }
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc_rt(smem_u32(tmem_slot), 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (i_begin >= i_end) {
    // (an empty q split: nothing to add to dK / dV / dQ)
  } else if (warp == 16) {
    // ---------------------------------------------------------------- TMA producer (warp-uniform)
    if (elect_one()) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      tma_prefetch_desc(&tmdO);
      mbar_expect_tx(smem_u32(bar_kv), 2 * TILE);
      tma_load_4d(smem_u32(sK), &tmK, smem_u32(bar_kv), 0, h, kv0, b);
      tma_load_4d(smem_u32(sV), &tmV, smem_u32(bar_kv), 0, h, kv0, b);
    }
    __syncwarp();
    for (int i = i_begin; i < i_end; ++i) {
      const int it = i - i_begin;
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      mbar_wait(smem_u32(&q_empty[s]), ph ^ 1);
      if (elect_one()) {
        const uint32_t fb = smem_u32(&q_full[s]);
        mbar_expect_tx(fb, 2 * TILE);
        tma_load_4d(smem_u32(sQ + s * TILE), &tmQ, fb, 0, h, i * 128, b);
        tma_load_4d(smem_u32(sdO + s * TILE), &tmdO, fb, 0, h, i * 128, b);
      }
      __syncwarp();
    }
  } else if (warp == 17) {
    // ---------------------------------------------------------------- MMA issuer (warp-uniform)
    const uint32_t idesc_kk = umma_idesc_f16(128, 128, 0, 0);    // S^T, dP^T : both operands K-major
    const uint32_t idesc_acc = umma_idesc_f16(128, p.dn, 0, 1);  // dV, dK    : B MN-major
    const uint32_t idesc_dq = umma_idesc_f16(128, p.dn, 1, 1);   // dQ = dS K : A (dS^T tile) and B MN-major
    const int nks = p.dn / 16;
    const uint32_t hi = umma_desc_hi_sw128(1024);
    const uint32_t k_lo = umma_desc_lo(smem_u32(sK), 16), v_lo = umma_desc_lo(smem_u32(sV), 16);
    const uint32_t k_mn = umma_desc_lo(smem_u32(sK), ABOX);
    const uint32_t pt_lo = umma_desc_lo(smem_u32(sPT), 16), ds_lo = umma_desc_lo(smem_u32(sDS), 16);
    const uint32_t ds_mn = umma_desc_lo(smem_u32(sDS), ABOX);
    auto issue_s = [&](int s) {  // S^T = K Q^T
      const uint32_t q_lo = umma_desc_lo(smem_u32(sQ + s * TILE), 16);
      for (int ks = 0; ks < nks; ++ks) {
        const uint32_t off = ((ks & 3) * 32) >> 4;
        umma_f16_ss(tmem + S_COL, umma_desc_pack(k_lo + off, hi), umma_desc_pack(q_lo + off, hi), idesc_kk, ks > 0);
      }
      umma_commit(smem_u32(bar_s));
    };
    auto issue_dp = [&](int s) {  // dP^T = V dO^T
      const uint32_t do_lo = umma_desc_lo(smem_u32(sdO + s * TILE), 16);
      for (int ks = 0; ks < nks; ++ks) {
        const uint32_t off = ((ks & 3) * 32) >> 4;
        umma_f16_ss(tmem + DP_COL, umma_desc_pack(v_lo + off, hi), umma_desc_pack(do_lo + off, hi), idesc_kk, ks > 0);
      }
      umma_commit(smem_u32(bar_dp));
    };
    mbar_wait(smem_u32(bar_kv), 0);
    mbar_wait(smem_u32(&q_full[0]), 0);
    tc_fence_after();
    if (elect_one()) {
      issue_s(0);
      issue_dp(0);
    }
    __syncwarp();
    for (int i = i_begin; i < i_end; ++i) {
      const int it = i - i_begin;
      const int s = it % STAGES;
      const uint32_t par = it & 1;
      const bool more = i + 1 < i_end;
      const int s1 = (it + 1) % STAGES;
      const uint32_t q_mn = umma_desc_lo(smem_u32(sQ + s * TILE), ABOX);
      const uint32_t do_mn = umma_desc_lo(smem_u32(sdO + s * TILE), ABOX);
      if (more) {  // (a) S^T of the next pair, as soon as this pair's scores have been read
        mbar_wait(smem_u32(&q_full[s1]), ((it + 1) / STAGES) & 1);
        mbar_wait(smem_u32(bar_sfree), par);
        tc_fence_after();
        if (elect_one()) issue_s(s1);
        __syncwarp();
      }
      // (b) dV += P^T dO
      mbar_wait(smem_u32(bar_pt), par);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t offp = ((ks >> 2) * ABOX + (ks & 3) * 32) >> 4;
          umma_f16_ss(tmem + DV_COL, umma_desc_pack(pt_lo + offp, hi), umma_desc_pack(do_mn + ks * 128, hi),
                      idesc_acc, (it > 0) || (ks > 0));
        }
        umma_commit(smem_u32(bar_dv));
      }
      __syncwarp();
      if (more) {  // (c) dP^T of the next pair, as soon as this pair's dP^T has been read
        mbar_wait(smem_u32(bar_dpfree), par);
        tc_fence_after();
        if (elect_one()) issue_dp(s1);
        __syncwarp();
      }
      // (d) dK += dS^T Q ; dQ = dS K (own accumulator: the previous pair's dQ must have been drained)
      mbar_wait(smem_u32(bar_ds), par);
      if (it > 0) mbar_wait(smem_u32(bar_dqfree), par ^ 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t offp = ((ks >> 2) * ABOX + (ks & 3) * 32) >> 4;
          umma_f16_ss(tmem + DK_COL, umma_desc_pack(ds_lo + offp, hi), umma_desc_pack(q_mn + ks * 128, hi),
                      idesc_acc, (it > 0) || (ks > 0));
        }
        if (p.dQacc) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma_f16_ss(tmem + DQ_COL, umma_desc_pack(ds_mn + ks * 128, hi), umma_desc_pack(k_mn + ks * 128, hi),
                        idesc_dq, ks > 0);
        }
        umma_commit(smem_u32(bar_dq));
        umma_commit(smem_u32(&q_empty[s]));
      }
      __syncwarp();
    }
  } else {
    // ---------------------------------------------------------------- softmax / dS / drain warps
    const int quarter = warp & 3, cg = warp >> 2;  // TMEM lane quarter; 32-column group of the q tile
    const int row = quarter * 32 + lane;           // kv row inside the tile (q row for the dQ drain)
    const int cb = cg * 32;
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    const bool kv_ok = kv0 + row < p.Nk;
    const long long bh = (long long)b * p.heads + h;
    auto load_ld = [&](int q) -> float {
      if (cg == 0) return q < p.Nq ? p.lse[bh * p.Nq + q] : INFINITY;  // +inf: exp2(. - L) == 0 for padding queries
      if (cg == 1) return q < p.Nq ? p.delta[bh * p.Nq + q] : 0.f;
      return 0.f;
    };
    // dQ(i) leaves through shared memory and ONE TMA reduce-add per pair (whole 128-byte lines added at the L2)
    // instead of 12 red.global.add.v4.f32 per thread: the reds alone were ~1600 cycles of each ~5000-cycle pair.
    auto drain_dq = [&](int i_prev, uint32_t par_prev) {
      mbar_wait(smem_u32(bar_dq), par_prev);  // dK / dQ of that pair retired (also: dS^T may be overwritten)
      tc_fence_after();
      if (p.dQacc && cg * 16 < p.dn) {
        const int c = cg * 16;
        uint32_t r[16];
        tmem_ld16(lane_addr + DQ_COL + c, r);
        tmem_ld_wait();
        float* dst = sDQ + row * p.d + c;
#pragma unroll
        for (int g = 0; g < 4; ++g)
          if (c + g * 4 < p.d)
            *reinterpret_cast<float4*>(dst + g * 4) =
                make_float4(__uint_as_float(r[g * 4 + 0]) * p.scale, __uint_as_float(r[g * 4 + 1]) * p.scale,
                            __uint_as_float(r[g * 4 + 2]) * p.scale, __uint_as_float(r[g * 4 + 3]) * p.scale);
      }
      tc_fence_before();
      mbar_arrive(smem_u32(bar_dqfree));  // the dQ accumulator is free for the next pair
      if (p.dQacc) {
        fence_async_smem();
        named_bar_sync(2, NSM);
        if (threadIdx.x == 0) {
          tma_reduce_add_3d(&tmDQ, smem_u32(sDQ), h * p.d, i_prev * 128, b);  // rows >= Nq are clipped by the map
          tma_store_commit();
        }
      }
    };

    float ld_next = load_ld(i_begin * 128 + row);
    for (int i = i_begin; i < i_end; ++i) {
      const int it = i - i_begin;
      const uint32_t par = it & 1;
      float* sLb = sL + par * 128;
      float* sDb = sD + par * 128;
      if (cg == 0) sLb[row] = ld_next;
      else if (cg == 1) sDb[row] = ld_next;
      if (i + 1 < i_end) ld_next = load_ld((i + 1) * 128 + row);
      if (threadIdx.x == 0) tma_store_wait_read<0>();  // the previous reduce-add has read the staging tile
      named_bar_sync(1, NSM);

      // ---- P phase: P^T = exp2(S^T * c - L)
      mbar_wait(smem_u32(bar_s), par);
      tc_fence_after();
      uint32_t pt[16];  // this thread's 32 entries of P^T as packed fp16, kept for the dS product
      {
        uint32_t r[32];
        tmem_ld32(lane_addr + S_COL + cb, r);
        tmem_ld_wait32(r);
        tc_fence_before();
        mbar_arrive(smem_u32(bar_sfree));  // S^T(i) is in registers: the next pair's scores may overwrite it
        if (it > 0) mbar_wait(smem_u32(bar_dv), par ^ 1);  // dV(i-1) has finished reading the P^T tile
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float4 l0 = *reinterpret_cast<const float4*>(sLb + cb + g * 8);
          const float4 l1 = *reinterpret_cast<const float4*>(sLb + cb + g * 8 + 4);
          uint32_t* o = pt + g * 4;
          o[0] = exp2_pack_mix<0>(fmaf(__uint_as_float(r[g * 8 + 0]), p.scale_log2, -l0.x),
                                  fmaf(__uint_as_float(r[g * 8 + 1]), p.scale_log2, -l0.y));
          o[1] = exp2_pack_mix<1>(fmaf(__uint_as_float(r[g * 8 + 2]), p.scale_log2, -l0.z),
                                  fmaf(__uint_as_float(r[g * 8 + 3]), p.scale_log2, -l0.w));
          o[2] = exp2_pack_mix<2>(fmaf(__uint_as_float(r[g * 8 + 4]), p.scale_log2, -l1.x),
                                  fmaf(__uint_as_float(r[g * 8 + 5]), p.scale_log2, -l1.y));
          o[3] = exp2_pack_mix<3>(fmaf(__uint_as_float(r[g * 8 + 6]), p.scale_log2, -l1.z),
                                  fmaf(__uint_as_float(r[g * 8 + 7]), p.scale_log2, -l1.w));
          if (!kv_ok) o[0] = o[1] = o[2] = o[3] = 0u;
          *reinterpret_cast<uint4*>(sPT + sw128_off(row, cg * 4 + g)) = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
      fence_async_smem();
      mbar_arrive(smem_u32(bar_pt));

      // ---- drain dQ of the previous pair (its MMA finished long ago)
      if (it > 0) drain_dq(i - 1, par ^ 1);

      // ---- dS phase: dS^T = P^T o (dP^T - delta)
      mbar_wait(smem_u32(bar_dp), par);
      tc_fence_after();
      {
        uint32_t r[32];
        tmem_ld32(lane_addr + DP_COL + cb, r);
        tmem_ld_wait32(r);
        tc_fence_before();
        mbar_arrive(smem_u32(bar_dpfree));  // dP^T(i) is in registers
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float4 d0 = *reinterpret_cast<const float4*>(sDb + cb + g * 8);
          const float4 d1 = *reinterpret_cast<const float4*>(sDb + cb + g * 8 + 4);
          const uint32_t* pp = pt + g * 4;
          uint4 o;
          o.x = hmul2_u32(pp[0], pack_half2(__uint_as_float(r[g * 8 + 0]) - d0.x, __uint_as_float(r[g * 8 + 1]) - d0.y));
          o.y = hmul2_u32(pp[1], pack_half2(__uint_as_float(r[g * 8 + 2]) - d0.z, __uint_as_float(r[g * 8 + 3]) - d0.w));
          o.z = hmul2_u32(pp[2], pack_half2(__uint_as_float(r[g * 8 + 4]) - d1.x, __uint_as_float(r[g * 8 + 5]) - d1.y));
          o.w = hmul2_u32(pp[3], pack_half2(__uint_as_float(r[g * 8 + 6]) - d1.z, __uint_as_float(r[g * 8 + 7]) - d1.w));
          *reinterpret_cast<uint4*>(sDS + sw128_off(row, cg * 4 + g)) = o;
        }
      }
      fence_async_smem();
      mbar_arrive(smem_u32(bar_ds));
    }
    // the last pair's dQ; its bar_dq also covers every MMA the CTA has issued (dV, dK accumulators are final)
    if (threadIdx.x == 0) tma_store_wait_read<0>();
    named_bar_sync(1, NSM);
    drain_dq(i_end - 1, (uint32_t)((i_end - 1 - i_begin) & 1));
    if (threadIdx.x == 0) tma_store_wait_all<0>();  // every reduce-add has been performed before the CTA retires

    // ------------------------------------------------------------ dV / dK epilogue (16 columns per warp group)
    const int c = cg * 16;
    if (c < p.dn) {
      uint32_t rv[16], rk[16];
      tmem_ld16(lane_addr + DV_COL + c, rv);
      tmem_ld16(lane_addr + DK_COL + c, rk);
      tmem_ld_wait();
      if (kv_ok && p.qsplit > 1) {
        const long long Cc = (long long)p.heads * p.d;
        float* wv = p.dkv_ws + ((long long)b * p.Nk + kv0 + row) * Cc + h * p.d;
        float* wk = wv + (long long)gridDim.z * p.Nk * Cc;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (c + g * 4 < p.d) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(wv + c + g * 4),
                         "f"(__uint_as_float(rv[g * 4 + 0])), "f"(__uint_as_float(rv[g * 4 + 1])),
                         "f"(__uint_as_float(rv[g * 4 + 2])), "f"(__uint_as_float(rv[g * 4 + 3]))
                         : "memory");
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(wk + c + g * 4),
                         "f"(__uint_as_float(rk[g * 4 + 0]) * p.scale), "f"(__uint_as_float(rk[g * 4 + 1]) * p.scale),
                         "f"(__uint_as_float(rk[g * 4 + 2]) * p.scale), "f"(__uint_as_float(rk[g * 4 + 3]) * p.scale)
                         : "memory");
          }
        }
      } else if (kv_ok) {
        tb::half_t* dv = p.dV + ((long long)b * p.Nk + kv0 + row) * p.lddv + h * p.d;
        tb::half_t* dk = p.dK + ((long long)b * p.Nk + kv0 + row) * p.lddk + h * p.d;
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
  if (warp == 0) tmem_dealloc_rt(tmem, 512u);
}

// ------------------------------------------------------------------------------------ backward, v3
// head_dim <= 64, non-causal (every UNet attention at 64x64 / 96x96, and the cross attentions).  Same loop as
// attn_bwd2_kernel -- one CTA per SM owns a 128-key KV tile of one (b, h), walks the query tiles, keeps dK / dV in
// TMEM and reduce-adds dQ tile by tile -- but with the lessons of the forward rewrite (profiles/r02_*):
//   * scores are computed as S = Q_i K^T / dP = dO_i V^T (rows = queries), not transposed: a thread owns a query row,
//     so the row's log-sum-exp L and delta are two REGISTERS per pair instead of 16 shared-memory vector loads per
//     thread per pair, and scale-and-subtract is one packed FFMA2 per two scores with broadcast scalars;
//     P and dS go to shared memory as [q][kv] tiles, read by the dV / dK MMAs as MN-major A operands (P^T, dS^T)
//     and by the dQ MMA as a K-major one;
//   * TWO MMA-issuing warps: one for the score MMAs (S, dP: they feed the softmax warps), one for the gradient MMAs
//     (dV, dK, dQ: they consume P and dS).  A single issuing warp pays ~100 cycles per mbarrier wake-up and per issue
//     burst and was the bottleneck of the forward kernel at 22 MMAs per iteration; a pair here is 30;
//   * no CTA-wide named barrier in the loop: the dQ tile leaves per 32-row slab (the four warps of a TMEM lane
//     quarter: 128-thread barrier, one TMA reduce-add each);
//   * all shared-memory traffic of the softmax warps is explicit st.shared (the generic stores of v2 resolved the
//     address space at run time).
// TMEM: S [0,128)  dP [128,256)  dQ [256,320)  dV [320,384)  dK [384,448).
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st_shared_v4f(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ uint32_t hmul2_sub(uint32_t pp, uint64_t dp2, uint64_t delta2) {
  // P (packed fp16 pair) * (dP - delta) (fp32 pair, rounded to fp16)
  float a, b;
  unpack_f32x2(sub_f32x2(dp2, delta2), a, b);
  uint32_t h;
  asm("cvt.rn." TB_H16X2 ".f32 %0, %1, %2;" : "=r"(h) : "f"(b), "f"(a));
  return hmul2_u32(pp, h);
}

template <int STAGES>
__global__ void __launch_bounds__(608, 1)
attn_bwd3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                 const __grid_constant__ CUtensorMap tmDQ, const AttnParams p) {
  constexpr int TILE = ABOX;
  constexpr uint32_t S_COL = 0, DP_COL = 128, DQ_COL = 256, DV_COL = 320, DK_COL = 384;
  constexpr int NSM = 512;  // softmax threads

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + TILE;
  uint8_t* sQ = sV + TILE;
  uint8_t* sdO = sQ + STAGES * TILE;
  uint8_t* sP = sdO + STAGES * TILE;   // P   [128 q x 128 kv] fp16, two swizzled 64-column boxes
  uint8_t* sDS = sP + 2 * ABOX;        // dS, same layout
  float* sDQ = reinterpret_cast<float*>(sDS + 2 * ABOX);  // 2 KB per softmax warp: fp32 staging of the dQ reduce-adds
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDQ + 128 * 64);
  uint64_t* bar_kv = bars;
  uint64_t* bar_s = bars + 1;       // S(i) complete                           (MMA-A -> softmax)
  uint64_t* bar_sfree = bars + 2;   // S(i) is in registers                    (softmax -> MMA-A)
  uint64_t* bar_p = bars + 3;       // P(i) in shared memory                   (softmax -> MMA-B)
  uint64_t* bar_dv = bars + 4;      // dV(i) has read P(i)                     (MMA-B -> softmax)
  uint64_t* bar_dp = bars + 5;      // dP(i) complete                          (MMA-A -> softmax)
  uint64_t* bar_dpfree = bars + 6;  // dP(i) is in registers                   (softmax -> MMA-A)
  uint64_t* bar_ds = bars + 7;      // dS(i) in shared memory                  (softmax -> MMA-B)
  uint64_t* bar_dq = bars + 8;      // dK(i), dQ(i) complete: dS(i) read       (MMA-B -> softmax)
  uint64_t* bar_dqfree = bars + 9;  // dQ(i) drained                           (softmax -> MMA-B)
  uint64_t* q_full = bars + 10;
  uint64_t* q_empty = bars + 10 + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10 + 2 * STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kv0 = (blockIdx.x / p.qsplit) * 128, h = blockIdx.y, b = blockIdx.z;
  const int qs = blockIdx.x % p.qsplit;
  const int q_per = (p.n_inner + p.qsplit - 1) / p.qsplit;
  const int i_begin = qs * q_per;
  const int i_end = min(p.n_inner, i_begin + q_per);
  const int n_it = i_end - i_begin;

  if (threadIdx.x == NSM) {
    mbar_init(smem_u32(bar_kv), 1);
    mbar_init(smem_u32(bar_s), 1);
    mbar_init(smem_u32(bar_sfree), NSM);
    mbar_init(smem_u32(bar_p), NSM);
    mbar_init(smem_u32(bar_dv), 1);
    mbar_init(smem_u32(bar_dp), 1);
    mbar_init(smem_u32(bar_dpfree), NSM);
    mbar_init(smem_u32(bar_ds), NSM);
    mbar_init(smem_u32(bar_dq), 1);
    mbar_init(smem_u32(bar_dqfree), NSM);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&q_full[s]), 1);
      mbar_init(smem_u32(&q_empty[s]), 1);
    }
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc_rt(smem_u32(tmem_slot), 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t hi = umma_desc_hi_sw128(1024);
  const int nks = p.dn / 16;

  if (n_it <= 0) {
    // (an empty q split: nothing to add to dK / dV / dQ)
  } else if (warp == 16) {
    // ---------------------------------------------------------------- TMA producer
    if (elect_one()) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      tma_prefetch_desc(&tmdO);
      mbar_expect_tx(smem_u32(bar_kv), 2 * TILE);
      tma_load_4d(smem_u32(sK), &tmK, smem_u32(bar_kv), 0, h, kv0, b);
      tma_load_4d(smem_u32(sV), &tmV, smem_u32(bar_kv), 0, h, kv0, b);
    }
    __syncwarp();
    for (int it = 0; it < n_it; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      mbar_wait(smem_u32(&q_empty[s]), ph ^ 1);
      if (elect_one()) {
        const uint32_t fb = smem_u32(&q_full[s]);
        mbar_expect_tx(fb, 2 * TILE);
        tma_load_4d(smem_u32(sQ + s * TILE), &tmQ, fb, 0, h, (i_begin + it) * 128, b);
        tma_load_4d(smem_u32(sdO + s * TILE), &tmdO, fb, 0, h, (i_begin + it) * 128, b);
      }
      __syncwarp();
    }
  } else if (warp == 17) {
    // ---------------------------------------------------------------- MMA-A: S = Q_i K^T, dP = dO_i V^T
    const uint32_t idesc_kk = umma_idesc_f16(128, 128, 0, 0);  // both operands K-major
    const uint32_t k_lo = umma_desc_lo(smem_u32(sK), 16), v_lo = umma_desc_lo(smem_u32(sV), 16);
    const uint32_t q_lo0 = umma_desc_lo(smem_u32(sQ), 16), do_lo0 = umma_desc_lo(smem_u32(sdO), 16);
    auto issue = [&](uint32_t d_tm, uint32_t a_lo, uint32_t b_lo, uint64_t* bar) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        if (ks < nks) umma_f16_ss(d_tm, umma_desc_pack(a_lo + ks * 2, hi), umma_desc_pack(b_lo + ks * 2, hi), idesc_kk, ks > 0);
      umma_commit(smem_u32(bar));
    };
    mbar_wait(smem_u32(bar_kv), 0);
    mbar_wait(smem_u32(&q_full[0]), 0);
    tc_fence_after();
    if (elect_one()) {
      issue(tmem + S_COL, q_lo0, k_lo, bar_s);
      issue(tmem + DP_COL, do_lo0, v_lo, bar_dp);
    }
    __syncwarp();
    for (int it0 = 0; it0 + 1 < n_it; it0 += STAGES) {
#pragma unroll
      for (int s = 0; s < STAGES; ++s) {
        const int it = it0 + s;
        if (it + 1 >= n_it) break;
        const int s1 = (s + 1) % STAGES;
        const uint32_t ph1 = ((it + 1) / STAGES) & 1;
        const uint32_t par = it & 1;
        mbar_wait(smem_u32(&q_full[s1]), ph1);
        TB_TRACE3(it, 0);
        mbar_wait(smem_u32(bar_sfree), par);  // S(it) has been read: the next scores may overwrite it
        tc_fence_after();
        TB_TRACE3(it, 1);
        if (elect_one()) issue(tmem + S_COL, q_lo0 + s1 * (TILE >> 4), k_lo, bar_s);
        __syncwarp();
        TB_TRACE3(it, 2);
        mbar_wait(smem_u32(bar_dpfree), par);
        tc_fence_after();
        TB_TRACE3(it, 3);
        if (elect_one()) issue(tmem + DP_COL, do_lo0 + s1 * (TILE >> 4), v_lo, bar_dp);
        __syncwarp();
        TB_TRACE3(it, 4);
      }
    }
  } else if (warp == 18) {
    // ---------------------------------------------------------------- MMA-B: dV += P^T dO, dK += dS^T Q, dQ = dS K
    // The softmax warps publish dS(i) and P(i+1) together, so per pair this warp issues dK(i), dQ(i), dV(i+1) back to
    // back (24 MMAs) after one pair of waits.
    const uint32_t idesc_acc = umma_idesc_f16(128, p.dn, 1, 1);  // A = P / dS read MN-major (transposed), B MN-major
    const uint32_t idesc_dq = umma_idesc_f16(128, p.dn, 0, 1);   // A = dS K-major, B = K MN-major
    const uint32_t p_mn = umma_desc_lo(smem_u32(sP), ABOX), ds_mn = umma_desc_lo(smem_u32(sDS), ABOX);
    const uint32_t ds_lo = umma_desc_lo(smem_u32(sDS), 16);
    const uint32_t k_mn = umma_desc_lo(smem_u32(sK), ABOX);
    const uint32_t q_mn0 = umma_desc_lo(smem_u32(sQ), ABOX), do_mn0 = umma_desc_lo(smem_u32(sdO), ABOX);
    auto issue_dv = [&](int it, uint32_t do_mn) {
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)
        umma_f16_ss(tmem + DV_COL, umma_desc_pack(p_mn + ks * 128, hi), umma_desc_pack(do_mn + ks * 128, hi), idesc_acc,
                    (it > 0) || (ks > 0));
      umma_commit(smem_u32(bar_dv));
    };
    mbar_wait(smem_u32(&q_full[0]), 0);
    mbar_wait(smem_u32(bar_p), 0);
    tc_fence_after();
    if (elect_one()) issue_dv(0, do_mn0);
    __syncwarp();
    for (int it0 = 0; it0 < n_it; it0 += STAGES) {
#pragma unroll
      for (int s = 0; s < STAGES; ++s) {
        const int it = it0 + s;
        if (it >= n_it) break;
        const int s1 = (s + 1) % STAGES;
        const uint32_t par = it & 1;
        const uint32_t q_mn = q_mn0 + s * (TILE >> 4);
        TB_TRACE3(it, 0);
        mbar_wait(smem_u32(bar_ds), par);
        if (it > 0) mbar_wait(smem_u32(bar_dqfree), par ^ 1);  // the previous pair's dQ has been drained
        tc_fence_after();
        TB_TRACE3(it, 1);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma_f16_ss(tmem + DK_COL, umma_desc_pack(ds_mn + ks * 128, hi), umma_desc_pack(q_mn + ks * 128, hi),
                        idesc_acc, (it > 0) || (ks > 0));
          if (p.dQacc || p.dQ16) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              const uint32_t off = ((ks >> 2) * ABOX + (ks & 3) * 32) >> 4;
              umma_f16_ss(tmem + DQ_COL, umma_desc_pack(ds_lo + off, hi), umma_desc_pack(k_mn + ks * 128, hi), idesc_dq,
                          ks > 0);
            }
          }
          umma_commit(smem_u32(bar_dq));
          umma_commit(smem_u32(&q_empty[s]));
        }
        __syncwarp();
        TB_TRACE3(it, 2);
        if (it + 1 < n_it) {
          mbar_wait(smem_u32(&q_full[s1]), ((it + 1) / STAGES) & 1);  // (long complete)
          mbar_wait(smem_u32(bar_p), par ^ 1);
          tc_fence_after();
          TB_TRACE3(it, 3);
          if (elect_one()) issue_dv(it + 1, do_mn0 + s1 * (TILE >> 4));
          __syncwarp();
          TB_TRACE3(it, 4);
        }
      }
    }
  } else if (warp < 16) {
    // ---------------------------------------------------------------- softmax / dS / drain warps
    // four threads per query row: warp w -> TMEM lane quarter w & 3, 32-column group w >> 2
    const int quarter = warp & 3, cg = warp >> 2;
    const int row = quarter * 32 + lane;  // query row inside the tile; kv row for the dK / dV epilogue
    const int cb = cg * 32;
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    const long long bh = (long long)b * p.heads + h;
    const uint32_t sp_base = smem_u32(sP), sds_base = smem_u32(sDS);
    const uint64_t sc2 = pack_f32x2(p.scale_log2, p.scale_log2);
    const int kvalid = p.Nk - kv0;  // columns >= kvalid of this KV tile are padding keys
    // Per iteration `it` a thread finishes pair it (dS(it) = P(it) o (dP(it) - delta)), starts pair it+1
    // (P(it+1) = exp2(S(it+1) c - L)) and moves its share of dQ(it-1) out: three independent instruction streams
    // (ALU / FMA, MUFU-bound, memory) in ONE pass with one proxy fence and no barrier wider than a warp, so the four
    // warps of an SM sub-partition drift apart and overlap each other's MUFU and ALU phases.  All waits on the
    // gradient MMAs (dS tile free, P tile free, dQ ready) sit at the end of the pass, a whole pass after their MMAs
    // were issued.  (First version of this kernel: P phase, dQ drain, dS phase one after the other with every warp in
    // lockstep: 3570 cycles per pair, of which 1130 MUFU-bound, 780 in the drain's barriers / fences, 450 dS arithmetic.)
    float l_cur = 0.f, d_cur = 0.f, l_nxt = 0.f, d_nxt = 0.f;
    auto load_ld = [&](int i, float& l, float& dl) {
      const int q = i * 128 + row;
      // +inf makes exp2(. - L) == 0 for padding queries
      l = q < p.Nq ? p.lse[bh * p.Nq + q] : INFINITY;
      dl = q < p.Nq ? p.delta[bh * p.Nq + q] : 0.f;
    };
    uint32_t pk[16];  // this thread's 32 entries of P(it) as packed fp16, kept for the dS product
    auto p_compute = [&](int it, float L, auto masked_tag) {  // P(it) -> pk
      constexpr bool MASKED = decltype(masked_tag)::value;
      const uint64_t negl2 = pack_f32x2(-L, -L);
      mbar_wait(smem_u32(bar_s), it & 1);
      tc_fence_after();
      uint32_t r[32];
      tmem_ld32(lane_addr + S_COL + cb, r);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(bar_sfree));  // S(it) is in registers: the next pair's scores may overwrite it
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const uint64_t e = fma_f32x2(pack_f32x2(__uint_as_float(r[2 * k]), __uint_as_float(r[2 * k + 1])), sc2, negl2);
        float p0, p1;
        exp2_pair_mufu(e, p0, p1);
        asm("cvt.rn." TB_H16X2 ".f32 %0, %1, %2;" : "=r"(pk[k]) : "f"(p1), "f"(p0));
        if (MASKED) {
          if (cb + 2 * k >= kvalid) pk[k] &= 0xffff0000u;
          if (cb + 2 * k + 1 >= kvalid) pk[k] &= 0x0000ffffu;
        }
      }
    };
    auto p_store = [&](int it) {
      if (it > 0) mbar_wait(smem_u32(bar_dv), (it - 1) & 1);  // dV(it-1) has finished reading the P tile
#pragma unroll
      for (int g = 0; g < 4; ++g)
        st_shared_v4(sp_base + sw128_off(row, cg * 4 + g), pk[g * 4], pk[g * 4 + 1], pk[g * 4 + 2], pk[g * 4 + 3]);
    };
    // dQ(it): each warp owns a [32 rows x 16 columns] block of the tile: TMEM -> its own 2 KB of staging -> (after the
    // proxy fence) two TMA reduce-adds of [32 x 8] floats.  Only __syncwarp is needed: lane 0 issued the previous
    // reduce-adds of this block and waits until they have read the staging block before the warp overwrites it.
    const bool dq_warp = (p.dQacc != nullptr || p.dQ16 != nullptr) && cg * 16 < p.dn;
    const uint32_t sdq_warp = smem_u32(sDQ) + (uint32_t)warp * 2048;  // [2 column groups][32 rows][8 floats]
    auto dq_to_staging = [&](int it) {
      mbar_wait(smem_u32(bar_dq), it & 1);  // dK / dQ of that pair retired (also: the dS tile may be overwritten)
      tc_fence_after();
      if (dq_warp && p.dQ16) {
        // single KV tile (cross attention): the tile's dQ is complete -> fp16 rows straight to global memory
        uint32_t r[16];
        tmem_ld16(lane_addr + DQ_COL + cg * 16, r);
        tmem_ld_wait();
        const int q = (i_begin + it) * 128 + row;
        if (q < p.Nq) {
          tb::half_t* dq = p.dQ16 + ((long long)b * p.Nq + q) * p.lddq16 + h * p.d + cg * 16;
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            if (cg * 16 + g * 8 < p.d) {
              uint4 o;
              o.x = pack_half2(__uint_as_float(r[g * 8 + 0]) * p.scale, __uint_as_float(r[g * 8 + 1]) * p.scale);
              o.y = pack_half2(__uint_as_float(r[g * 8 + 2]) * p.scale, __uint_as_float(r[g * 8 + 3]) * p.scale);
              o.z = pack_half2(__uint_as_float(r[g * 8 + 4]) * p.scale, __uint_as_float(r[g * 8 + 5]) * p.scale);
              o.w = pack_half2(__uint_as_float(r[g * 8 + 6]) * p.scale, __uint_as_float(r[g * 8 + 7]) * p.scale);
              *reinterpret_cast<uint4*>(dq + g * 8) = o;
            }
          }
        }
      } else if (dq_warp) {
        uint32_t r[16];
        tmem_ld16(lane_addr + DQ_COL + cg * 16, r);
        tmem_ld_wait();
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const uint32_t dst = sdq_warp + (uint32_t)(g * 32 + lane) * 32;
          st_shared_v4f(dst, __uint_as_float(r[g * 8 + 0]) * p.scale, __uint_as_float(r[g * 8 + 1]) * p.scale,
                        __uint_as_float(r[g * 8 + 2]) * p.scale, __uint_as_float(r[g * 8 + 3]) * p.scale);
          st_shared_v4f(dst + 16, __uint_as_float(r[g * 8 + 4]) * p.scale, __uint_as_float(r[g * 8 + 5]) * p.scale,
                        __uint_as_float(r[g * 8 + 6]) * p.scale, __uint_as_float(r[g * 8 + 7]) * p.scale);
        }
      }
      tc_fence_before();
      mbar_arrive(smem_u32(bar_dqfree));
    };
    auto dq_reduce = [&](int it) {  // after fence.proxy.async
      if (!dq_warp || p.dQ16) return;
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int g = 0; g < 2; ++g)
          if (cg * 16 + g * 8 < p.d)  // (rows >= Nq are clipped by the tensor map)
            tma_reduce_add_3d(&tmDQ, sdq_warp + g * 1024, h * p.d + cg * 16 + g * 8, (i_begin + it) * 128 + quarter * 32, b);
        tma_store_commit();
      }
    };
    auto body = [&](auto masked_tag) {
      load_ld(i_begin, l_cur, d_cur);
      if (n_it > 1) load_ld(i_begin + 1, l_nxt, d_nxt);
      p_compute(0, l_cur, masked_tag);
      p_store(0);
      fence_async_smem();
      mbar_arrive(smem_u32(bar_p));
      for (int it = 0; it < n_it; ++it) {
        const uint32_t par = it & 1;
        const bool more = it + 1 < n_it;
        const uint64_t delta2 = pack_f32x2(d_cur, d_cur);
        // ---- dS(it) = P(it) o (dP(it) - delta), kept in registers
        TB_TRACE3(it, 1);
        mbar_wait(smem_u32(bar_dp), par);
        tc_fence_after();
        TB_TRACE3(it, 2);
        uint32_t ds[16];
        {
          uint32_t r[32];
          tmem_ld32(lane_addr + DP_COL + cb, r);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(smem_u32(bar_dpfree));  // dP(it) is in registers
#pragma unroll
          for (int k = 0; k < 16; ++k)
            ds[k] = hmul2_sub(pk[k], pack_f32x2(__uint_as_float(r[2 * k]), __uint_as_float(r[2 * k + 1])), delta2);
        }
        TB_TRACE3(it, 3);
        // ---- P(it+1) -> pk
        if (more) {
          l_cur = l_nxt;
          d_cur = d_nxt;
          if (it + 2 < n_it) load_ld(i_begin + it + 2, l_nxt, d_nxt);
          p_compute(it + 1, l_cur, masked_tag);
        }
        TB_TRACE3(it, 4);
        // ---- everything that waits for the gradient MMAs of the previous pass
        if (it > 0) dq_to_staging(it - 1);  // (also: dK / dQ(it-1) have finished reading the dS tile)
#pragma unroll
        for (int g = 0; g < 4; ++g)
          st_shared_v4(sds_base + sw128_off(row, cg * 4 + g), ds[g * 4], ds[g * 4 + 1], ds[g * 4 + 2], ds[g * 4 + 3]);
        if (more) p_store(it + 1);
        fence_async_smem();
        mbar_arrive(smem_u32(bar_ds));
        if (more) mbar_arrive(smem_u32(bar_p));
        if (it > 0) dq_reduce(it - 1);
        TB_TRACE3(it, 5);
      }
      // the last pair's dQ; its bar_dq also covers every MMA the gradient warp has issued (dV, dK are final)
      dq_to_staging(n_it - 1);
      fence_async_smem();
      dq_reduce(n_it - 1);
      if (lane == 0) tma_store_wait_all<0>();  // every reduce-add has been performed before the CTA retires
    };
    if (kvalid >= 128) body(std::false_type{});
    else body(std::true_type{});

    // ------------------------------------------------------------ dV / dK epilogue (16 columns per warp group)
    const bool kv_ok = kv0 + row < p.Nk;
    const int c = cg * 16;
    if (c < p.dn) {
      uint32_t rv[16], rk[16];
      tmem_ld16(lane_addr + DV_COL + c, rv);
      tmem_ld16(lane_addr + DK_COL + c, rk);
      tmem_ld_wait();
      if (kv_ok && p.qsplit > 1) {
        const long long Cc = (long long)p.heads * p.d;
        float* wv = p.dkv_ws + ((long long)b * p.Nk + kv0 + row) * Cc + h * p.d;
        float* wk = wv + (long long)gridDim.z * p.Nk * Cc;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (c + g * 4 < p.d) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(wv + c + g * 4),
                         "f"(__uint_as_float(rv[g * 4 + 0])), "f"(__uint_as_float(rv[g * 4 + 1])),
                         "f"(__uint_as_float(rv[g * 4 + 2])), "f"(__uint_as_float(rv[g * 4 + 3]))
                         : "memory");
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(wk + c + g * 4),
                         "f"(__uint_as_float(rk[g * 4 + 0]) * p.scale), "f"(__uint_as_float(rk[g * 4 + 1]) * p.scale),
                         "f"(__uint_as_float(rk[g * 4 + 2]) * p.scale), "f"(__uint_as_float(rk[g * 4 + 3]) * p.scale)
                         : "memory");
          }
        }
      } else if (kv_ok) {
        tb::half_t* dv = p.dV + ((long long)b * p.Nk + kv0 + row) * p.lddv + h * p.d;
        tb::half_t* dk = p.dK + ((long long)b * p.Nk + kv0 + row) * p.lddk + h * p.d;
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
  if (warp == 0) tmem_dealloc_rt(tmem, 512u);
}

}  // namespace tb

// ------------------------------------------------------------------------------------ host side
using namespace tb;

static long long* g_attn_trace = nullptr;
// debug hook (not part of the reference-facing surface): timestamps of CTA (0,0,0) of the next forward launches
extern "C" int tb_attn_debug_trace(void* buf) {
  g_attn_trace = (long long*)buf;
  return TB_OK;
}

static int make_head_map(CUtensorMap* m, const void* base, long long ld, int B, int heads, int N,
                         int d) {
  uint64_t dims[4] = {(uint64_t)d, (uint64_t)heads, (uint64_t)N, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)d * 2, (uint64_t)ld * 2, (uint64_t)N * ld * 2};
  uint32_t box[4] = {64, 1, 128, 1};
  return make_tmap_f16(m, base, 4, dims, strides, box);
}

static int check_attn_args(const char* fn, int B, int heads, int Nq, int Nk, int d, long long ldq,
                           long long ldk, long long ldv, int causal) {
  TB_REQUIRE(B > 0 && heads > 0 && Nq > 0 && Nk > 0, TB_E_SHAPE, "%s: bad sizes", fn);
  TB_REQUIRE(d % 8 == 0 && d >= 8 && d <= 192, TB_E_SHAPE, "%s: head_dim %d unsupported (d %% 8 == 0, d <= 192)", fn, d);
  TB_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0, TB_E_ALIGN, "%s: row strides must be multiples of 8", fn);
  TB_REQUIRE(ldq >= (long long)heads * d && ldk >= (long long)heads * d && ldv >= (long long)heads * d,
             TB_E_SHAPE, "%s: row stride smaller than heads*d", fn);
  TB_REQUIRE(causal == 0 || causal == 1, TB_E_ARG, "%s: causal must be 0 or 1", fn);
  TB_REQUIRE(!causal || Nq == Nk, TB_E_SHAPE, "%s: causal attention needs Nq == Nk", fn);
  return TB_OK;
}

template <int NB, int STAGES>
static int launch_attn_fwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                           const AttnParams& p, int B, cudaStream_t st) {
  constexpr int smem = NB * ABOX * (1 + 2 * STAGES) + ABOX + 1024 + 256 + 1024;
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
  attn_fwd_kernel<NB, STAGES><<<grid, 320, smem, st>>>(tq, tk, tv, p);
  return check_launch("attn_fwd_kernel");
}

template <int NB, int STAGES>
static int launch_attn_fwd2(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                            const AttnParams& p, int B, cudaStream_t st) {
  constexpr int smem = NB * ABOX * (2 + 2 * STAGES) + 512 + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd2_kernel<NB, STAGES>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(attn_fwd2<%d,%d>, %d): %s", NB, STAGES, smem, cudaGetErrorString(e));
      return TB_E_CUDA;
    }
    configured = true;
  }
  dim3 grid((p.Nq + 255) / 256, p.heads, B);
  attn_fwd2_kernel<NB, STAGES><<<grid, 384, smem, st>>>(tq, tk, tv, p);
  return check_launch("attn_fwd2_kernel");
}

template <int NB, int STAGES>
static int launch_attn_fwd2h(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                             const AttnParams& p, int B, cudaStream_t st) {
  constexpr int smem = NB * ABOX * (2 + 2 * STAGES) + 4096 + 512 + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd2h_kernel<NB, STAGES>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(attn_fwd2h<%d,%d>, %d): %s", NB, STAGES, smem, cudaGetErrorString(e));
      return TB_E_CUDA;
    }
    configured = true;
  }
  dim3 grid((p.Nq + 255) / 256, p.heads, B);
  attn_fwd2h_kernel<NB, STAGES><<<grid, 608, smem, st>>>(tq, tk, tv, p);
  return check_launch("attn_fwd2h_kernel");
}

static int launch_attn_bwd2(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                            const CUtensorMap& tdo, const AttnParams& p, int B, cudaStream_t st) {
  constexpr int STAGES = 2;
  constexpr int smem = ABOX * (2 + 2 * STAGES) + 4 * ABOX + 128 * 64 * 4 + 2048 + 256 + 1024;
  CUtensorMap tdq = tq;  // (unused without dQ)
  if (p.dQacc) {
    uint64_t dims[3] = {(uint64_t)p.lddq, (uint64_t)p.Nq, (uint64_t)B};
    uint64_t strides[2] = {(uint64_t)p.lddq * 4, (uint64_t)p.Nq * p.lddq * 4};
    uint32_t box[3] = {(uint32_t)p.d, 128u, 1u};
    int rc = make_tmap_f32_plain(&tdq, p.dQacc, 3, dims, strides, box);
    if (rc) return rc;
  }
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd2_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(attn_bwd2, %d): %s", smem, cudaGetErrorString(e));
      return TB_E_CUDA;
    }
    configured = true;
  }
  dim3 grid(((p.Nk + 127) / 128) * p.qsplit, p.heads, B);
  attn_bwd2_kernel<STAGES><<<grid, 576, smem, st>>>(tq, tk, tv, tdo, tdq, p);
  return check_launch("attn_bwd2_kernel");
}

static int launch_attn_bwd3(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                            const CUtensorMap& tdo, const AttnParams& p, int B, cudaStream_t st) {
  constexpr int STAGES = 3;
  constexpr int smem = ABOX * (2 + 2 * STAGES) + 4 * ABOX + 128 * 64 * 4 + 256 + 1024;
  CUtensorMap tdq = tq;  // (unused without dQ)
  if (p.dQacc) {
    uint64_t dims[3] = {(uint64_t)p.lddq, (uint64_t)p.Nq, (uint64_t)B};
    uint64_t strides[2] = {(uint64_t)p.lddq * 4, (uint64_t)p.Nq * p.lddq * 4};
    uint32_t box[3] = {8u, 32u, 1u};  // one [32 rows x 8 columns] block per reduce-add (two per softmax warp)
    int rc = make_tmap_f32_plain(&tdq, p.dQacc, 3, dims, strides, box);
    if (rc) return rc;
  }
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd3_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(attn_bwd3, %d): %s", smem, cudaGetErrorString(e));
      return TB_E_CUDA;
    }
    configured = true;
  }
  dim3 grid(((p.Nk + 127) / 128) * p.qsplit, p.heads, B);
  attn_bwd3_kernel<STAGES><<<grid, 608, smem, st>>>(tq, tk, tv, tdo, tdq, p);
  return check_launch("attn_bwd3_kernel");
}

template <int NB, int STAGES>
static int launch_attn_bwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                           const CUtensorMap& tdo, const AttnParams& p, int B, cudaStream_t st) {
  constexpr int smem = NB * ABOX * (2 + 2 * STAGES) + 2 * ABOX + 1024 + 256 + 1024;
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
  dim3 grid(((p.Nk + 127) / 128) * p.qsplit, p.heads, B);
  attn_bwd_kernel<NB, STAGES><<<grid, 320, smem, st>>>(tq, tk, tv, tdo, p);
  return check_launch("attn_bwd_kernel");
}

extern "C" int tb_attn_fwd_f16(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                               int64_t ldv, void* o, int64_t ldo, float* lse, int B, int heads, int Nq,
                               int Nk, int d, float scale, int causal, void* stream) {
  int rc = tb_check_device();
  if (rc) return rc;
  TB_REQUIRE(q && k && v && o, TB_E_ARG, "tb_attn_fwd_f16: null pointer");
  rc = check_attn_args("tb_attn_fwd_f16", B, heads, Nq, Nk, d, ldq, ldk, ldv, causal);
  if (rc) return rc;
  TB_REQUIRE(ldo % 8 == 0, TB_E_ALIGN, "tb_attn_fwd_f16: ldo %% 8");
  CUtensorMap tq, tk, tv;
  if ((rc = make_head_map(&tq, q, ldq, B, heads, Nq, d))) return rc;
  if ((rc = make_head_map(&tk, k, ldk, B, heads, Nk, d))) return rc;
  if ((rc = make_head_map(&tv, v, ldv, B, heads, Nk, d))) return rc;
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.Nq = Nq; p.Nk = Nk; p.heads = heads; p.d = d; p.dn = (d + 15) / 16 * 16;
  p.causal = causal;
  p.tmem_cols = (128 + p.dn + 16 <= 256) ? 256 : 512;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.O = (tb::half_t*)o; p.ldo = ldo; p.lse = lse;
  p.trace = g_attn_trace;
  p.n_inner = (Nk + 127) / 128;
  const int nb = (d + 63) / 64;
  cudaStream_t st = (cudaStream_t)stream;
  // long non-causal sequences: the two-query-tile ping-pong kernel (one CTA per SM)
  static const bool v1 = getenv("TB_ATTN_FWD_V1") != nullptr;  // diagnostic switch: the two-CTA-per-SM kernel
  if (!v1 && !causal && p.dn <= 128 && Nq >= 256 && Nk >= 256) {
    static const bool one_thread = getenv("TB_ATTN_FWD2_ONE_THREAD_PER_ROW") != nullptr;  // diagnostic switch
    if (one_thread) {
      if (nb == 1) return launch_attn_fwd2<1, 4>(tq, tk, tv, p, B, st);
      return launch_attn_fwd2<2, 2>(tq, tk, tv, p, B, st);
    }
    if (nb == 1) return launch_attn_fwd2h<1, 4>(tq, tk, tv, p, B, st);
    return launch_attn_fwd2h<2, 2>(tq, tk, tv, p, B, st);
  }
  if (nb == 1) return launch_attn_fwd<1, 2>(tq, tk, tv, p, B, st);
  if (nb == 2) return launch_attn_fwd<2, 1>(tq, tk, tv, p, B, st);
  return launch_attn_fwd<3, 1>(tq, tk, tv, p, B, st);
}

extern "C" int tb_attn_bwd_f16(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                               int64_t ldv, const void* o, int64_t ldo, const void* dO, int64_t lddo,
                               const float* lse, float* delta, float* dQacc, int64_t lddq, void* dQ16,
                               int64_t lddq16, void* dK, int64_t lddk, void* dV, int64_t lddv, int B, int heads,
                               int Nq, int Nk, int d, float scale, int causal, void* stream) {
  int rc = tb_check_device();
  if (rc) return rc;
  TB_REQUIRE(q && k && v && o && dO && lse && delta && dK && dV, TB_E_ARG, "tb_attn_bwd_f16: null pointer");
  rc = check_attn_args("tb_attn_bwd_f16", B, heads, Nq, Nk, d, ldq, ldk, ldv, causal);
  if (rc) return rc;
  TB_REQUIRE(ldo % 8 == 0 && lddo % 8 == 0 && lddk % 8 == 0 && lddv % 8 == 0 && lddq % 4 == 0,
             TB_E_ALIGN, "tb_attn_bwd_f16: stride alignment");
  TB_REQUIRE(heads * d <= 1280, TB_E_SHAPE, "tb_attn_bwd_f16: heads*d = %d > 1280", heads * d);
  // dQ16: with a single KV tile every dQ row is complete inside one CTA, so it is stored once as fp16 and the fp32
  // accumulator (its memset, the reduce-adds and the cast that followed) is not needed
  TB_REQUIRE(!dQ16 || (!dQacc && Nk <= 128 && lddq16 % 8 == 0 && d % 8 == 0), TB_E_ARG,
             "tb_attn_bwd_f16: dQ16 needs Nk <= 128 (Nk=%d), no dQacc, lddq16 %% 8 == 0", Nk);
  cudaStream_t st = (cudaStream_t)stream;
  {
    const long long rows = (long long)B * Nq;
    const int wpb = 8;
    attn_delta_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, st>>>(
        (const tb::half_t*)o, (const tb::half_t*)dO, ldo, lddo, delta, B, Nq, heads, d);
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
  p.causal = causal;
  p.tmem_cols = (128 + 2 * p.dn <= 256) ? 256 : 512;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.lse = const_cast<float*>(lse);
  p.delta = delta;
  p.dQacc = dQacc; p.lddq = lddq;
  p.dQ16 = (tb::half_t*)dQ16; p.lddq16 = lddq16;
  p.dK = (tb::half_t*)dK; p.lddk = lddk;
  p.dV = (tb::half_t*)dV; p.lddv = lddv;
  p.n_inner = (Nq + 127) / 128;
  p.trace = g_attn_trace;
  {
    static const char* ex = getenv("TB_ATTN_EXPERIMENT");
    p.experiment = ex ? atoi(ex) : 0;
  }
  const int nb = (d + 63) / 64;
  // under-filled grid (cross attention: a single KV tile per (b, h)) -> split the Q loop over several CTAs
  p.qsplit = 1;
  p.dkv_ws = nullptr;
  {
    static const bool off = getenv("TB_ATTN_NO_QSPLIT") != nullptr;  // diagnostic switch
    const int ctas = ((Nk + 127) / 128) * heads * B;
    static const bool v1q = getenv("TB_ATTN_BWD_V1") != nullptr;
    const int per_sm = (nb == 1 && (causal || v1q)) ? 2 : 1;  // resident CTAs per SM
    const Workspace* w = find_ws(stream);
    const size_t need = 2ull * B * Nk * heads * d * sizeof(float);
    if (!off && !causal && w && ctas * 2 <= num_sms() * per_sm && p.n_inner >= 4 && (heads * d) % 4 == 0 &&
        WS_COUNTER_BYTES + need <= w->bytes && lddk % 4 == 0 && lddv % 4 == 0) {
      int qsp = (num_sms() * per_sm) / ctas;
      if (qsp > p.n_inner / 2) qsp = p.n_inner / 2;
      if (qsp >= 2) {
        const int q_per = (p.n_inner + qsp - 1) / qsp;
        qsp = (p.n_inner + q_per - 1) / q_per;  // no empty split (its CTA would add an unwritten accumulator)
      }
      if (qsp >= 2) {
        p.qsplit = qsp;
        p.dkv_ws = reinterpret_cast<float*>(w->base + WS_COUNTER_BYTES);
        cudaError_t e = cudaMemsetAsync(p.dkv_ws, 0, need, st);
        TB_REQUIRE(e == cudaSuccess, TB_E_CUDA, "tb_attn_bwd_f16: memset dkv_ws: %s", cudaGetErrorString(e));
      }
    }
  }
  static const bool v1 = getenv("TB_ATTN_BWD_V1") != nullptr;  // diagnostic switch: the two-CTA-per-SM kernel
  static const bool v2 = getenv("TB_ATTN_BWD_V2") != nullptr;  // diagnostic switch: the transposed-score kernel
  if (nb == 1 && !causal && !v1 && !v2) rc = launch_attn_bwd3(tq, tk, tv, tdo, p, B, st);
  else if (nb == 1 && !causal && !v1 && !dQ16) rc = launch_attn_bwd2(tq, tk, tv, tdo, p, B, st);
  else if (nb == 1) rc = launch_attn_bwd<1, 1>(tq, tk, tv, tdo, p, B, st);
  else if (nb == 2) rc = launch_attn_bwd<2, 1>(tq, tk, tv, tdo, p, B, st);
  else rc = launch_attn_bwd<3, 1>(tq, tk, tv, tdo, p, B, st);
  if (rc || p.qsplit == 1) return rc;
  const long long rows = (long long)B * Nk;
  const long long nvec = 2 * rows * (heads * d / 4);
  attn_dkv_finish_kernel<<<(unsigned)((nvec + 255) / 256 > 1184 ? 1184 : (nvec + 255) / 256), 256, 0, st>>>(
      p.dkv_ws, (tb::half_t*)dK, lddk, (tb::half_t*)dV, lddv, rows, heads * d);
  return check_launch("attn_dkv_finish_kernel");
}
