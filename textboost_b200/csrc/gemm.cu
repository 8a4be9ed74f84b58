// tcgen05 / TMEM / TMA GEMM with fused epilogue, and conv3x3 as an implicit GEMM on the same
// mainloop (A operand gathered by a 4-D TMA box over the NHWC tensor, borders zero-filled by TMA).
//
//   C[M,N] = epilogue( A[M,K] * B[N,K]^T )      A, B fp16 K-major, accumulate fp32 in TMEM.
//
// CTA = 320 threads: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane issues
// tcgen05.mma), warps 2..9 = epilogue (tcgen05.ld -> registers -> fused bias/act/residual -> smem box ->
// TMA store), two warps per TMEM lane quarter, software-pipelined over 32-column chunks.
// PERSISTENT: one CTA per SM walks the 128 x BN output tiles (n fastest, so concurrently running CTAs
// share the A tile through L2).  Three pipelines run across tile boundaries: the STAGES-deep smem ring
// (TMA <-> MMA, full/empty mbarriers), a two-deep TMEM accumulator ring (MMA <-> epilogue,
// tmem_full/tmem_empty) so the epilogue of tile i overlaps the mainloop of tile i+1, and the tile walk.
#include <stdlib.h>

#include "host_util.h"
#include "sm100.cuh"

namespace tb {

struct GemmParams {
  int M, N, K;        // logical sizes (K = 9*Cin for conv)
  int num_kb;         // number of 64-wide k blocks
  int n_tiles;        // ceil(N / BN)
  int m_tiles;        // ceil(M / 128)
  // conv geometry (CONV only)
  int H, W, kb_per_tap;
  // rows of an output tile that hold pixels: 128, or 96 for the 96/48/24/12-wide levels of a 768^2 image, where
  // no 128-pixel box tiles the plane (the MMA still runs M = 128; rows >= tile_rows are never loaded or stored)
  int tile_rows;
  int a_stage_bytes;  // bytes one TMA box of the A operand delivers (tile_rows x 128)
  // MMA M extent: 128, or 64 for under-filled problems on the generic epilogue (the 616-row text-encoder GEMMs with
  // N = 768: 60 tiles of 128 rows leave 88 SMs idle while each CTA walks its K loop at ~0.2 us per k-block; 120 tiles
  // of 64 rows halve the operand bytes per CTA and k-block).  With M = 64 the accumulator rows sit in TMEM lanes
  // 0-15 of each 32-lane quarter (row r -> lane 32*(r/16) + r%16): lanes 16-31 of the epilogue warps idle.
  int mma_m;          // (layout found with scripts/m64_diag.py: the other candidate, lanes 0-63, reads garbage)
  // epilogue
  void* C;
  long long ldc;
  const tb::half_t* bias;
  const tb::half_t* rowvec;
  long long ldrv;     // row stride of rowvec (>= N: a column slice of one batched time-embedding projection)
  int rows_per_group;
  const void* residual;
  long long ldr;
  int res_f32;
  float alpha;
  int act;
  int out_kind;
  int tma_store;      // fp16 output through shared memory + TMA tile stores (full 128-byte lines)
  // split-K (under-filled problems): a work item = (tile, split); every split parks its fp32 partial tile in `ws`,
  // bumps counters[tile], and the last arriver adds the partials in split order and runs the epilogue
  int splits;         // 1 = off
  int kb_per_split;
  int experiment;     // timing experiments only (TB_GEMM_EXPERIMENT): 1 = skip the B-operand loads, 2 = skip the A loads
  float* ws;          // [num_tiles * splits][128][BN] fp32
  int* counters;      // [num_tiles], zero between launches (the fixing CTA resets its tile's counter)
};

// Deliberately NOT inlined: the epilogue is unrolled 16x8 elements and an inlined three-way activation (erff
// alone is ~35 instructions) blew the kernel up to ~14k SASS instructions, whose instruction-cache misses
// ("no_instructions" was the top stall reason in ncu) cost more than the activation ever does.  Only the
// time-embedding MLP uses a GEMM-fused activation on the hot path.
__device__ __noinline__ float apply_act(float x, int act) {
  switch (act) {
    case TB_ACT_SILU: return x / (1.f + __expf(-x));
    case TB_ACT_QUICK_GELU: return x / (1.f + __expf(-1.702f * x));
    case TB_ACT_GELU: return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
    default: return x;
  }
}

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KiB
constexpr int C_BOX_BYTES = BM * 64 * 2;    // 16 KiB

template <int BN>
constexpr int tmem_cols() {
  return BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
}

template <int BN, int STAGES, bool CONV, bool FAST>
__global__ void __launch_bounds__(320, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                         const __grid_constant__ CUtensorMap tmB,
                                                         const __grid_constant__ CUtensorMap tmC,
                                                         const GemmParams p) {
  constexpr int B_STAGE_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  constexpr int ACC_STRIDE = tmem_cols<BN>();       // columns per accumulator stage
  constexpr int TMEM_COLS = 2 * ACC_STRIDE;         // two accumulator stages

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // manual 1 KiB alignment (SWIZZLE_128B atoms)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sC = smem + STAGES * STAGE_BYTES;  // two [128 x 64] fp16 staging boxes for the TMA tile stores
  uint64_t* bars = reinterpret_cast<uint64_t*>(sC + 2 * C_BOX_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full_bar = bars + 2 * STAGES;       // [2]
  uint64_t* tmem_empty_bar = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  volatile uint32_t* split_flag = tmem_slot + 1;
  // FAST: bias[n] + rowvec[image, n] of the current tile (16-byte aligned: read as float4)
  float* sBV = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 2) + 15) & ~static_cast<uintptr_t>(15));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_work = p.m_tiles * p.n_tiles * p.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.tma_store) tma_prefetch_desc(&tmC);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tmem_full_bar[a]), 1);
      mbar_init(smem_u32(&tmem_empty_bar[a]), 256);
    }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------ TMA producer (warp-uniform; one elected lane issues)
    int stage = 0;
    uint32_t phase = 0;
    for (int work = blockIdx.x; work < num_work; work += gridDim.x) {
      const int tile = work / p.splits;
      const int kb0 = (work - tile * p.splits) * p.kb_per_split;
      const int kb1 = min(kb0 + p.kb_per_split, p.num_kb);
      const int n_tile = tile % p.n_tiles;
      const int m_tile = tile / p.n_tiles;
      int cw = 0, ch = 0, cb = 0;
      if (CONV) {
        const int p0 = m_tile * p.tile_rows;
        cw = p0 % p.W;
        ch = (p0 / p.W) % p.H;
        cb = p0 / (p.W * p.H);
      }
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
        if (elect_one()) {
          const uint32_t fb = smem_u32(&full_bar[stage]);
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sb = sa + A_STAGE_BYTES;
          mbar_expect_tx(fb, (p.experiment == 2 ? 0 : p.a_stage_bytes) + (p.experiment == 1 ? 0 : B_STAGE_BYTES));
          if (p.experiment == 2) {
          } else if (CONV) {
            const int tap = kb / p.kb_per_tap;
            const int c0 = (kb - tap * p.kb_per_tap) * BK;
            const int ky = tap / 3, kx = tap - ky * 3;
            tma_load_4d(sa, &tmA, fb, c0, cw + kx - 1, ch + ky - 1, cb);
          } else {
            tma_load_2d(sa, &tmA, fb, kb * BK, m_tile * p.tile_rows);
          }
          if (p.experiment != 1) tma_load_2d(sb, &tmB, fb, kb * BK, n_tile * BN);
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (warp-uniform; one elected lane issues)
    const uint32_t idesc = umma_idesc_f16(p.mma_m, BN, 0, 0);
    const uint32_t dhi = umma_desc_hi_sw128(1024);
    int stage = 0;
    uint32_t phase = 0;
    int lt = 0;
    for (int work = blockIdx.x; work < num_work; work += gridDim.x, ++lt) {
      const int kb0 = (work % p.splits) * p.kb_per_split;
      const int kb1 = min(kb0 + p.kb_per_split, p.num_kb);
      const int acc = lt & 1;
      mbar_wait(smem_u32(&tmem_empty_bar[acc]), ((lt >> 1) & 1) ^ 1);  // epilogue drained this stage
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * ACC_STRIDE;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
        const uint32_t alo = umma_desc_lo(sa, 16), blo = umma_desc_lo(sa + A_STAGE_BYTES, 16);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_f16_ss(d_tmem, umma_desc_pack(alo + k * 2, dhi), umma_desc_pack(blo + k * 2, dhi), idesc,
                        ((kb - kb0) | k) != 0);
          umma_commit(smem_u32(&empty_bar[stage]));  // frees the smem slot when these MMAs retire
          if (kb == kb1 - 1) umma_commit(smem_u32(&tmem_full_bar[acc]));
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------------------------ epilogue (warps 2..9, two warps per TMEM lane quarter)
    // Warp w may touch TMEM lanes 32*(w%4)..+31.  The two warps of a quarter split every 64-column box:
    // half 0 takes columns [0,32), half 1 [32,64) of each box, so each SM sub-partition always has two
    // epilogue warps to interleave.  Per warp the chunk loop is software-pipelined: the tcgen05.ld of chunk
    // i+1 and the residual loads of chunk i+1 are in flight while chunk i is converted and stored.
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const bool m64 = !FAST && p.mma_m == 64;
    // TMEM lane of this thread is always 32*quarter + lane; with M = 64 that lane holds tile row 16*quarter + lane
    // (lanes 0-15 of the quarter) and lanes 16-31 hold nothing
    const bool lane_live = !m64 || lane < 16;
    const int row = !m64 ? quarter * 32 + lane : quarter * 16 + (lane & 15);
    const bool issuer = threadIdx.x == 64;
    constexpr int NCH = BN / 32;           // 32-column chunks per tile
    constexpr int MY_MAX = (NCH + 1) / 2;  // chunks per warp (half 0 takes the odd one out)
    int lt = 0;
    uint32_t n_box = 0;  // staging boxes written so far (selects the buffer)
    if constexpr (FAST) {
      // ---- FAST epilogue (the UNet's common case, chosen by the host): fp16 output through the TMA-store
      // boxes, alpha = 1, no activation, optional bias / per-image row vector / fp16 residual, M % 128 == 0,
      // N % BN == 0, a tile never straddles two images.  ~110 instructions per 32-column chunk instead of ~460:
      // bias + rowvec are summed once per tile into shared memory (fp32, broadcast reads), the accumulator is
      // rounded to fp16 once and the residual is added in half2 — the reference adds its residual to the
      // fp16-rounded conv / linear output the same way.
      const int epi_tid = threadIdx.x - 64;
      for (int work = blockIdx.x; work < num_work; work += gridDim.x, ++lt) {
        const int n_tile = work % p.n_tiles;
        const int m_tile = work / p.n_tiles;
        const int acc = lt & 1;
        const long long m = (long long)m_tile * p.tile_rows + row;
        const int ncol0 = n_tile * BN;
        const uint32_t acc_addr = tmem_base + acc * ACC_STRIDE + ((uint32_t)(quarter * 32) << 16);
        const tb::half_t* res = (p.residual && row < p.tile_rows)
                                ? reinterpret_cast<const tb::half_t*>(p.residual) + m * p.ldr + ncol0
                                : nullptr;
        const bool row_live = row < p.tile_rows;
        named_bar_sync(3, 256);  // every warp has finished reading the previous tile's sBV
        if (epi_tid < BN) {
          float bv = p.bias ? tb::h2f(p.bias[ncol0 + epi_tid]) : 0.f;
          if (p.rowvec)
            bv += tb::h2f(
                p.rowvec[((long long)m_tile * p.tile_rows / p.rows_per_group) * p.ldrv + ncol0 + epi_tid]);
          sBV[epi_tid] = bv;
        }
        uint4 rq[4], rq_next[4];
        if (res && half < NCH) {
#pragma unroll
          for (int j = 0; j < 4; ++j) rq[j] = *reinterpret_cast<const uint4*>(res + half * 32 + 8 * j);
        }
        named_bar_sync(3, 256);
        mbar_wait(smem_u32(&tmem_full_bar[acc]), (lt >> 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int i = 0; i < MY_MAX; ++i) {
          const int c = (2 * i + half) * 32;
          const bool have = c < BN;
          const bool staged = (c & ~63) + 64 <= BN;
          uint8_t* box = sC + (n_box & 1) * C_BOX_BYTES;
          if (have) {
            uint32_t r[32];
            tmem_ld32(acc_addr + c, r);
            if (res && c + 64 < BN) {
#pragma unroll
              for (int j = 0; j < 4; ++j) rq_next[j] = *reinterpret_cast<const uint4*>(res + c + 64 + 8 * j);
            }
            tmem_ld_wait32(r);
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              const float4 b0 = *reinterpret_cast<const float4*>(sBV + c + j);
              const float4 b1 = *reinterpret_cast<const float4*>(sBV + c + j + 4);
              uint4 o;
              o.x = pack_half2(__uint_as_float(r[j]) + b0.x, __uint_as_float(r[j + 1]) + b0.y);
              o.y = pack_half2(__uint_as_float(r[j + 2]) + b0.z, __uint_as_float(r[j + 3]) + b0.w);
              o.z = pack_half2(__uint_as_float(r[j + 4]) + b1.x, __uint_as_float(r[j + 5]) + b1.y);
              o.w = pack_half2(__uint_as_float(r[j + 6]) + b1.z, __uint_as_float(r[j + 7]) + b1.w);
              if (res) {
                const uint4 q = rq[j >> 3];
                *reinterpret_cast<tb::half2_t*>(&o.x) = __hadd2(*reinterpret_cast<const tb::half2_t*>(&o.x), *reinterpret_cast<const tb::half2_t*>(&q.x));
                *reinterpret_cast<tb::half2_t*>(&o.y) = __hadd2(*reinterpret_cast<const tb::half2_t*>(&o.y), *reinterpret_cast<const tb::half2_t*>(&q.y));
                *reinterpret_cast<tb::half2_t*>(&o.z) = __hadd2(*reinterpret_cast<const tb::half2_t*>(&o.z), *reinterpret_cast<const tb::half2_t*>(&q.z));
                *reinterpret_cast<tb::half2_t*>(&o.w) = __hadd2(*reinterpret_cast<const tb::half2_t*>(&o.w), *reinterpret_cast<const tb::half2_t*>(&q.w));
              }
              if (staged) {
                const int chunk = ((c & 63) + j) >> 3;
                *reinterpret_cast<uint4*>(box + row * 128 + ((chunk ^ (row & 7)) << 4)) = o;
              } else if (row_live) {
                *reinterpret_cast<uint4*>(reinterpret_cast<tb::half_t*>(p.C) + m * p.ldc + ncol0 + c + j) = o;
              }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) rq[j] = rq_next[j];
          }
          if (64 * i + 64 <= BN) {
            fence_async_smem();
            if (issuer) tma_store_wait_read<0>();
            named_bar_sync(1, 256);
            if (issuer) {
              tma_store_2d(&tmC, smem_u32(box), ncol0 + 64 * i, m_tile * p.tile_rows);
              tma_store_commit();
            }
            ++n_box;
          }
        }
        tc_fence_before();
        mbar_arrive(smem_u32(&tmem_empty_bar[acc]));
      }
    } else
    for (int work = blockIdx.x; work < num_work; work += gridDim.x, ++lt) {
      const int tile = work / p.splits;
      const int n_tile = tile % p.n_tiles;
      const int m_tile = tile / p.n_tiles;
      const int acc = lt & 1;
      const long long m = (long long)m_tile * p.tile_rows + row;
      bool from_ws = false;  // split-K fix-up: the accumulator chunks come from the fp32 workspace
      if (p.splits > 1) {
        // ---- park this split's partial tile, then find out whether we are the last split of the tile
        const uint32_t a_addr = tmem_base + acc * ACC_STRIDE + ((uint32_t)(quarter * 32) << 16);
        float* slot = p.ws + ((size_t)work * BM + row) * BN;
        mbar_wait(smem_u32(&tmem_full_bar[acc]), (lt >> 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c = half * 32; c < BN; c += 64) {
          uint32_t t[32];
          tmem_ld32(a_addr + c, t);
          tmem_ld_wait32(t);
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            __stcg(reinterpret_cast<float4*>(slot + c + j),
                   make_float4(__uint_as_float(t[j]), __uint_as_float(t[j + 1]), __uint_as_float(t[j + 2]),
                               __uint_as_float(t[j + 3])));
        }
        tc_fence_before();
        mbar_arrive(smem_u32(&tmem_empty_bar[acc]));  // the accumulator stage is free again
        __threadfence();
        named_bar_sync(2, 256);
        if (issuer) *split_flag = atomicAdd(p.counters + tile, 1) == p.splits - 1;
        named_bar_sync(2, 256);
        const bool last = *split_flag != 0;
        named_bar_sync(2, 256);  // everyone has read the flag before the next work item overwrites it
        if (!last) continue;
        __threadfence();
        if (issuer) p.counters[tile] = 0;  // ready for the next launch
        from_ws = true;
      }
      const uint32_t acc_addr = tmem_base + acc * ACC_STRIDE + ((uint32_t)(quarter * 32) << 16);
      const bool row_ok = m < p.M && row < p.tile_rows && lane_live;
      const tb::half_t* rv = nullptr;
      if (p.rowvec && row_ok) rv = p.rowvec + (m / p.rows_per_group) * p.ldrv;
      const tb::half_t* res = nullptr;
      const float* res32 = nullptr;
      if (p.residual && row_ok) {
        if (p.res_f32) res32 = reinterpret_cast<const float*>(p.residual) + m * p.ldr;
        else res = reinterpret_cast<const tb::half_t*>(p.residual) + m * p.ldr;
      }
      const int ncol0 = n_tile * BN;
      // The chunk loop is deliberately ROLLED (one copy of the ~600-instruction chunk body): fully unrolled and
      // software-pipelined, the epilogue was 11-14k SASS instructions and ncu's top stall reason was
      // "no_instructions" (instruction-cache misses on straight-line code run once per tile).  Latency is hidden
      // by the second epilogue warp of each sub-partition and by prefetching the next chunk's residual.
      uint32_t r[32];
      uint4 rq[4], rq_next[4];
      auto load_res = [&](int c, uint4* q) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int n = ncol0 + c + 8 * j;
          q[j] = (res && n < p.N) ? *reinterpret_cast<const uint4*>(res + n) : make_uint4(0, 0, 0, 0);
        }
      };
      // sum of the parked partials of one 32-column chunk, in split order (deterministic)
      const float* ws_row = p.ws + ((size_t)tile * p.splits * BM + row) * BN;
      if (half < NCH) load_res(half * 32, rq);  // residual of the first chunk: before the accumulator wait
      if (!from_ws) {
        mbar_wait(smem_u32(&tmem_full_bar[acc]), (lt >> 1) & 1);
        tc_fence_after();
      }

#pragma unroll 1
      for (int i = 0; i < MY_MAX; ++i) {
        const int c = (2 * i + half) * 32;  // first column of this warp's chunk inside the tile
        const int cn = c + 64;              // its next chunk
        const bool have = c < BN;
        const bool staged = p.tma_store && ((c & ~63) + 64 <= BN);  // full 64-column box: smem + TMA store
        uint8_t* box = sC + (n_box & 1) * C_BOX_BYTES;
        if (have) {
          if (from_ws) {
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = 0;
#pragma unroll 1
            for (int sp = 0; sp < p.splits; ++sp) {
              const float* src = ws_row + (size_t)sp * BM * BN + c;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 f = __ldcg(reinterpret_cast<const float4*>(src + j));
                r[j] = __float_as_uint(__uint_as_float(r[j]) + f.x);
                r[j + 1] = __float_as_uint(__uint_as_float(r[j + 1]) + f.y);
                r[j + 2] = __float_as_uint(__uint_as_float(r[j + 2]) + f.z);
                r[j + 3] = __float_as_uint(__uint_as_float(r[j + 3]) + f.w);
              }
            }
          } else {
            tmem_ld32(acc_addr + c, r);
          }
          if (cn < BN) load_res(cn, rq_next);
          // every global load of the chunk is issued here, before the TMEM wait and the arithmetic, so their
          // latencies overlap (left inside the per-group branches they were each exposed in turn)
          const int n0 = ncol0 + c;
          uint4 qb[4], qv[4];
          float4 q32[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int n = n0 + 8 * j;
            const bool ok = row_ok && n < p.N;
            qb[j] = (p.bias && ok) ? __ldg(reinterpret_cast<const uint4*>(p.bias + n)) : make_uint4(0, 0, 0, 0);
            qv[j] = (rv && ok) ? __ldg(reinterpret_cast<const uint4*>(rv + n)) : make_uint4(0, 0, 0, 0);
            if (res32 && ok) {
              q32[2 * j] = *reinterpret_cast<const float4*>(res32 + n);
              q32[2 * j + 1] = *reinterpret_cast<const float4*>(res32 + n + 4);
            } else {
              q32[2 * j] = q32[2 * j + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
          if (!from_ws) tmem_ld_wait32(r);
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            const int n = n0 + j;
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = __uint_as_float(r[j + k]);
            if (p.alpha != 1.f) {
#pragma unroll
              for (int k = 0; k < 8; ++k) v[k] *= p.alpha;
            }
            {
              const tb::half2_t* hb = reinterpret_cast<const tb::half2_t*>(&qb[j >> 3]);
              const tb::half2_t* hv = reinterpret_cast<const tb::half2_t*>(&qv[j >> 3]);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float2 fb = tb::h22f2(hb[k]), fv = tb::h22f2(hv[k]);
                v[2 * k] += fb.x + fv.x;
                v[2 * k + 1] += fb.y + fv.y;
              }
            }
            if (p.act != TB_ACT_NONE) {
#pragma unroll
              for (int k = 0; k < 8; ++k) v[k] = apply_act(v[k], p.act);
            }
            if (p.residual) {
              const tb::half2_t* h = reinterpret_cast<const tb::half2_t*>(&rq[j >> 3]);
              const float4 a0 = q32[2 * (j >> 3)], a1 = q32[2 * (j >> 3) + 1];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float2 f = tb::h22f2(h[k]);
                v[2 * k] += f.x;
                v[2 * k + 1] += f.y;
              }
              v[0] += a0.x; v[1] += a0.y; v[2] += a0.z; v[3] += a0.w;
              v[4] += a1.x; v[5] += a1.y; v[6] += a1.z; v[7] += a1.w;
            }
            if (staged) {
              uint4 o;
              o.x = pack_half2(v[0], v[1]);
              o.y = pack_half2(v[2], v[3]);
              o.z = pack_half2(v[4], v[5]);
              o.w = pack_half2(v[6], v[7]);
              const int chunk = ((c & 63) + j) >> 3;  // 16-byte chunk inside the 128-byte box row
              if (lane_live) *reinterpret_cast<uint4*>(box + row * 128 + ((chunk ^ (row & 7)) << 4)) = o;
            } else if (row_ok && n < p.N) {
              if (p.out_kind == TB_OUT_F16) {
                uint4 o;
                o.x = pack_half2(v[0], v[1]);
                o.y = pack_half2(v[2], v[3]);
                o.z = pack_half2(v[4], v[5]);
                o.w = pack_half2(v[6], v[7]);
                *reinterpret_cast<uint4*>(reinterpret_cast<tb::half_t*>(p.C) + m * p.ldc + n) = o;
              } else {
                float* cp = reinterpret_cast<float*>(p.C) + m * p.ldc + n;
                float4 o0 = make_float4(v[0], v[1], v[2], v[3]);
                float4 o1 = make_float4(v[4], v[5], v[6], v[7]);
                if (p.out_kind == TB_OUT_F32_ACC) {
                  const float4 a0 = *reinterpret_cast<const float4*>(cp);
                  const float4 a1 = *reinterpret_cast<const float4*>(cp + 4);
                  o0.x += a0.x; o0.y += a0.y; o0.z += a0.z; o0.w += a0.w;
                  o1.x += a1.x; o1.y += a1.y; o1.z += a1.z; o1.w += a1.w;
                }
                *reinterpret_cast<float4*>(cp) = o0;
                *reinterpret_cast<float4*>(cp + 4) = o1;
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) rq[j] = rq_next[j];
        }
        // box i of the tile (columns [64i, 64i+64)) is complete once both halves have written their chunk
        if (p.tma_store && 64 * i + 64 <= BN) {
          fence_async_smem();
          if (issuer) tma_store_wait_read<0>();  // the previous box's store has read its buffer out
          named_bar_sync(1, 256);
          if (issuer) {
            tma_store_2d(&tmC, smem_u32(box), ncol0 + 64 * i, m_tile * p.tile_rows);
            tma_store_commit();
          }
          ++n_box;
        }
      }
      if (!from_ws) {
        tc_fence_before();
        mbar_arrive(smem_u32(&tmem_empty_bar[acc]));  // 256 arrivals release the accumulator stage
      }
    }
    if (issuer) tma_store_wait_read<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------ host side

// Decide the split-K factor for an under-filled problem (tiles <= half the SMs and a long K loop).
static void plan_split(GemmParams& p, int bn, int tiles, cudaStream_t st) {
  p.splits = 1;
  p.kb_per_split = p.num_kb;
  p.ws = nullptr;
  p.counters = nullptr;
  static const bool off = getenv("TB_GEMM_NO_SPLITK") != nullptr;  // diagnostic switch
  const Workspace* w = find_ws((void*)st);
  // Measured (scripts/probe_small_gemm.py): the fix-up (partials to L2, __threadfence, counter round trip, the
  // last CTA re-reading splits x 64 KB) costs ~8-15 us, so splitting only pays for very long K loops -- the 3x3
  // convolutions of the 8x8 / 16x16 levels (180-360 k-blocks: 80 -> 32 us).  GEMMs with K <= 5120 lose.
  static const int min_kb = getenv("TB_GEMM_SPLITK_MIN_KB") ? atoi(getenv("TB_GEMM_SPLITK_MIN_KB")) : 128;  // tuning knob
  if (off || !w || tiles * 2 > num_sms() || p.num_kb < min_kb || p.mma_m == 64) return;
  int s = num_sms() / tiles;
  if (s > p.num_kb / 4) s = p.num_kb / 4;
  if (s > 16) s = 16;
  if (s < 2) return;
  const int kbps = (p.num_kb + s - 1) / s;
  s = (p.num_kb + kbps - 1) / kbps;  // every split owns at least one k block
  const size_t need = (size_t)tiles * s * BM * bn * sizeof(float);
  if (s < 2 || tiles > (int)(WS_COUNTER_BYTES / sizeof(int)) || WS_COUNTER_BYTES + need > w->bytes) return;
  p.splits = s;
  p.kb_per_split = kbps;
  p.counters = reinterpret_cast<int*>(w->base);
  p.ws = reinterpret_cast<float*>(w->base + WS_COUNTER_BYTES);
}

template <int BN, int STAGES>
constexpr int gemm_smem_bytes() {
  return STAGES * (A_STAGE_BYTES + BN * BK * 2) + 2 * C_BOX_BYTES + (2 * STAGES + 5) * 8 + 16 + 256 * 4 + 1024;
}

template <int BN, int STAGES, bool CONV, bool FAST>
static int launch_gemm_v(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                         GemmParams& p, int m_tiles, cudaStream_t st) {
  constexpr int smem = gemm_smem_bytes<BN, STAGES>();
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES, CONV, FAST>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(gemm<%d,%d>): %s", BN, STAGES, cudaGetErrorString(e));
      return TB_E_CUDA;
    }
    configured = true;
  }
  const int tiles = p.n_tiles * m_tiles * p.splits;
  dim3 grid(tiles < num_sms() ? tiles : num_sms());
  gemm_tc_kernel<BN, STAGES, CONV, FAST><<<grid, 320, smem, st>>>(tmA, tmB, tmC, p);
  return check_launch("gemm_tc_kernel");
}

template <int BN, int STAGES, bool CONV>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                       GemmParams& p, int m_tiles, cudaStream_t st) {
  p.n_tiles = (p.N + BN - 1) / BN;
  p.m_tiles = m_tiles;
  {
    static const int ex = getenv("TB_GEMM_EXPERIMENT") ? atoi(getenv("TB_GEMM_EXPERIMENT")) : 0;
    p.experiment = ex;
  }
  plan_split(p, BN, p.n_tiles * m_tiles, st);
  static const bool no_fast = getenv("TB_GEMM_NO_FAST_EPILOGUE") != nullptr;  // diagnostic switch
  const bool fast = BN >= 64 && !no_fast && p.mma_m == 128 && p.splits == 1 && p.tma_store == 1 && p.out_kind == TB_OUT_F16 &&
                    p.alpha == 1.f && p.act == TB_ACT_NONE && !p.res_f32 && p.N % BN == 0 && p.M % p.tile_rows == 0 &&
                    (!p.rowvec || p.rows_per_group % p.tile_rows == 0);
  if (fast) return launch_gemm_v<BN, STAGES, CONV, true>(tmA, tmB, tmC, p, m_tiles, st);
  return launch_gemm_v<BN, STAGES, CONV, false>(tmA, tmB, tmC, p, m_tiles, st);
}

// Choose the N tile.  Among the tile widths that divide N, minimise  waves x (128 + BN)  where waves =
// ceil(tiles / SMs): per k-block a CTA moves 128 + BN operand rows, and a persistent grid finishes when its busiest
// CTA does, so e.g. [2048,1280] takes 128 tiles of 160 (one wave) rather than 80 of 256, and [512,10240] takes 256
// tiles of 160 (two short rounds) rather than 160 of 256 (two long ones).  Ties go to the wider tile.  Problems
// that will be split along K (>= 128 k-blocks, under half a wave) stay on >= 128-wide tiles: there the split,
// not the tile, fills the machine.
static int pick_bn(int N, int m_tiles, int num_kb) {
  static const int cands[4] = {256, 160, 128, 64};
  const int sms = num_sms();
  static const int forced = getenv("TB_GEMM_BN") ? atoi(getenv("TB_GEMM_BN")) : 0;  // tuning knob
  if (forced && N % forced == 0) return forced;
  int best = 0;
  long long best_cost = 0;
  for (int i = 0; i < 4; ++i) {
    const int c = cands[i];
    if (N % c != 0) continue;
    const long long tiles = (long long)m_tiles * (N / c);
    if (num_kb >= 128 && c < 128 && m_tiles * ((N + 127) / 128) * 2 <= sms) continue;
    const long long cost = ((tiles + sms - 1) / sms) * (128 + c);
    if (best == 0 || cost < best_cost) {
      best = c;
      best_cost = cost;
    }
  }
  if (best) return best;
  if (N >= 1024) return 256;  // ragged N: the tensor map clips the last tile
  if (N > 96) return 128;
  if (N > 32) return 64;
  return 32;
}

template <bool CONV>
static int dispatch_gemm(const CUtensorMap& tmA, const void* Bw, long long ldb, GemmParams& p,
                         int m_tiles, cudaStream_t st) {
  const int bn = pick_bn(p.N, m_tiles, p.num_kb);
  CUtensorMap tmB;
  {
    uint64_t dims[2] = {(uint64_t)p.K, (uint64_t)p.N};
    uint64_t strides[1] = {(uint64_t)ldb * 2};
    uint32_t box[2] = {(uint32_t)BK, (uint32_t)bn};
    int rc = make_tmap_f16(&tmB, Bw, 2, dims, strides, box);
    if (rc) return rc;
  }
  // fp16 outputs leave through shared memory + TMA tile stores (box 64 columns x 128 rows)
  CUtensorMap tmC = tmB;
  p.tma_store = 0;
  static const bool direct_store = getenv("TB_GEMM_DIRECT_STORE") != nullptr;  // diagnostic switch
  if (p.out_kind == TB_OUT_F16 && bn >= 64 && !direct_store) {
    uint64_t dims[2] = {(uint64_t)p.N, (uint64_t)p.M};
    uint64_t strides[1] = {(uint64_t)p.ldc * 2};
    uint32_t box[2] = {64u, (uint32_t)p.tile_rows};
    int rc = make_tmap_f16(&tmC, p.C, 2, dims, strides, box);
    if (rc) return rc;
    p.tma_store = 1;
  }
  // (Deeper TMA rings for the under-filled cases -- 6 x 32 KB at BN = 128, 8 x 24 KB at BN = 64 -- were measured
  // and did not help: those launches are not bound by bytes in flight.)
  switch (bn) {
    case 256: return launch_gemm<256, 4, CONV>(tmA, tmB, tmC, p, m_tiles, st);
    case 160: return launch_gemm<160, 4, CONV>(tmA, tmB, tmC, p, m_tiles, st);
    case 128: return launch_gemm<128, 4, CONV>(tmA, tmB, tmC, p, m_tiles, st);
    case 64: return launch_gemm<64, 4, CONV>(tmA, tmB, tmC, p, m_tiles, st);
    default: return launch_gemm<32, 4, CONV>(tmA, tmB, tmC, p, m_tiles, st);
  }
}

static int fill_epilogue(GemmParams& p, void* C, long long ldc, const tb_epilogue* ep) {
  p.C = C;
  p.ldc = ldc;
  p.bias = nullptr;
  p.rowvec = nullptr;
  p.rows_per_group = 1;
  p.residual = nullptr;
  p.ldr = 0;
  p.res_f32 = 0;
  p.alpha = 1.f;
  p.act = TB_ACT_NONE;
  p.out_kind = TB_OUT_F16;
  if (ep) {
    p.bias = reinterpret_cast<const tb::half_t*>(ep->bias);
    p.rowvec = reinterpret_cast<const tb::half_t*>(ep->rowvec);
    p.rows_per_group = ep->rows_per_group > 0 ? ep->rows_per_group : 1;
    p.ldrv = ep->ld_rowvec > 0 ? ep->ld_rowvec : 0;
    p.residual = ep->residual;
    p.ldr = ep->ldr;
    p.res_f32 = ep->residual_f32;
    p.alpha = ep->alpha;
    p.act = ep->act;
    p.out_kind = ep->out_kind;
    TB_REQUIRE(p.act >= 0 && p.act <= 3, TB_E_ARG, "epilogue: unknown activation %d", p.act);
    TB_REQUIRE(p.out_kind >= 0 && p.out_kind <= 2, TB_E_ARG, "epilogue: unknown out_kind %d",
               p.out_kind);
    TB_REQUIRE(!p.residual || (p.ldr % 8 == 0), TB_E_ALIGN, "epilogue: ldr %% 8 != 0");
    TB_REQUIRE(((uintptr_t)p.bias | (uintptr_t)p.rowvec | (uintptr_t)p.residual) % 16 == 0,
               TB_E_ALIGN, "epilogue: bias/rowvec/residual must be 16-byte aligned");
  }
  if (p.ldrv == 0) p.ldrv = p.N;
  TB_REQUIRE(!p.rowvec || (p.ldrv >= p.N && p.ldrv % 8 == 0), TB_E_ALIGN,
             "epilogue: ld_rowvec must be >= N and a multiple of 8 (%lld)", p.ldrv);
  const int esz = p.out_kind == TB_OUT_F16 ? 2 : 4;
  TB_REQUIRE(((uintptr_t)C % 16 == 0) && ((ldc * esz) % 16 == 0), TB_E_ALIGN,
             "output pointer / ldc not 16-byte aligned");
  return TB_OK;
}

}  // namespace tb

using namespace tb;

extern "C" int tb_gemm_f16(const void* A, int64_t lda, const void* B, int64_t ldb, void* C,
                           int64_t ldc, int M, int N, int K, const tb_epilogue* ep, void* stream) {
  int rc = tb_check_device();
  if (rc) return rc;
  TB_REQUIRE(A && B && C, TB_E_ARG, "tb_gemm_f16: null pointer");
  TB_REQUIRE(M > 0 && N > 0 && K > 0, TB_E_SHAPE, "tb_gemm_f16: M,N,K must be positive (%d,%d,%d)",
             M, N, K);
  TB_REQUIRE(N % 8 == 0 && K % 8 == 0, TB_E_SHAPE, "tb_gemm_f16: N and K must be multiples of 8");
  TB_REQUIRE(lda % 8 == 0 && ldb % 8 == 0 && lda >= K && ldb >= K, TB_E_ALIGN,
             "tb_gemm_f16: lda/ldb must be >= K and multiples of 8");
  TB_REQUIRE(((uintptr_t)A | (uintptr_t)B) % 16 == 0, TB_E_ALIGN, "tb_gemm_f16: A/B alignment");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M;
  p.N = N;
  p.K = K;
  p.num_kb = (K + BK - 1) / BK;
  p.tile_rows = BM;
  p.a_stage_bytes = A_STAGE_BYTES;
  p.mma_m = 128;
  rc = fill_epilogue(p, C, ldc, ep);
  if (rc) return rc;
  {
    // 64-row tiles when 128-row tiles would fill less than half the machine and the problem is on the generic
    // epilogue anyway (ragged M, fp32 output / residual, activation): the text-encoder GEMMs with N = 768 / 784
    static const int m64 = getenv("TB_GEMM_M64") ? atoi(getenv("TB_GEMM_M64")) : 1;  // 0 = off (A/B switch)
    const int mt128 = (M + BM - 1) / BM;
    const int bn = pick_bn(N, mt128, p.num_kb);
    const bool generic = M % BM != 0 || p.out_kind != TB_OUT_F16 || p.res_f32 || p.act != TB_ACT_NONE || p.alpha != 1.f;
    if (m64 && generic && M > 64 && mt128 * ((N + bn - 1) / bn) * 2 <= num_sms() && p.num_kb < 128) {
      p.mma_m = 64;
      p.tile_rows = 64;
      p.a_stage_bytes = 64 * BK * 2;
    }
  }
  CUtensorMap tmA;
  uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
  uint64_t strides[1] = {(uint64_t)lda * 2};
  uint32_t box[2] = {(uint32_t)BK, (uint32_t)p.tile_rows};
  rc = make_tmap_f16(&tmA, A, 2, dims, strides, box);
  if (rc) return rc;
  return dispatch_gemm<false>(tmA, B, ldb, p, (M + p.tile_rows - 1) / p.tile_rows, (cudaStream_t)stream);
}

extern "C" int tb_conv3x3_f16(const void* x, const void* w, void* y, int B, int H, int W, int Cin,
                              int Cout, const tb_epilogue* ep, void* stream) {
  int rc = tb_check_device();
  if (rc) return rc;
  TB_REQUIRE(x && w && y, TB_E_ARG, "tb_conv3x3_f16: null pointer");
  TB_REQUIRE(B > 0 && H > 0 && W > 0, TB_E_SHAPE, "tb_conv3x3_f16: bad B/H/W");
  TB_REQUIRE(Cin % 64 == 0 && Cout % 8 == 0, TB_E_SHAPE,
             "tb_conv3x3_f16: Cin %% 64 and Cout %% 8 required (Cin=%d Cout=%d)", Cin, Cout);
  // An output tile = box_b images x box_h rows x box_w columns of pixels, contiguous in (b,h,w) order, as many as
  // fit in the 128 rows of the MMA: 128 at the power-of-two SD sizes, 96 at the 96 / 48 / 24-wide levels of a
  // 768^2 image, 72 at its 12x12 level.  rows always divides B*H*W.
  int bw, bh = 1, bb = 1;
  if (W >= 128) {
    bw = W % 128 == 0 ? 128 : 96;
    TB_REQUIRE(W % bw == 0, TB_E_SHAPE, "tb_conv3x3_f16: W=%d tiles neither 128 nor 96 pixels", W);
  } else {
    bw = W;
    for (int c = 1; c <= H && bw * c <= 128; ++c)
      if (H % c == 0) bh = c;
    if (bh == H)
      for (int c = 1; c <= B && bw * bh * c <= 128; ++c)
        if (B % c == 0) bb = c;
  }
  const int rows = bw * bh * bb;
  TB_REQUIRE(rows >= 1 && rows <= 128, TB_E_SHAPE, "tb_conv3x3_f16: %dx%d does not tile", H, W);
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = B * H * W;
  p.N = Cout;
  p.K = 9 * Cin;
  p.kb_per_tap = Cin / BK;
  p.num_kb = 9 * p.kb_per_tap;
  p.H = H;
  p.W = W;
  p.tile_rows = rows;
  p.a_stage_bytes = rows * BK * 2;
  p.mma_m = 128;
  rc = fill_epilogue(p, y, Cout, ep);
  if (rc) return rc;
  CUtensorMap tmA;
  uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
  uint32_t box[4] = {(uint32_t)BK, (uint32_t)bw, (uint32_t)bh, (uint32_t)bb};
  rc = make_tmap_f16(&tmA, x, 4, dims, strides, box);
  if (rc) return rc;
  return dispatch_gemm<true>(tmA, w, (long long)9 * Cin, p, (p.M + rows - 1) / rows,
                             (cudaStream_t)stream);
}
