// Microbenchmark (diagnostic): throughput of the pipes the attention softmax leans on, per SM:
// tcgen05.ld / tcgen05.st (TMEM <-> registers), MUFU.EX2, packed fp32x2 FMA, cvt.f16x2, FMNMX3 -- with 4, 8 or 16
// warps resident on one SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I textboost_b200/csrc
#include <cstdio>
#include "sm100.cuh"
using namespace tb;

enum { LD32 = 0, ST16 = 1, EX2 = 2, FFMA2 = 3, CVT = 4, MAX3 = 5, MIX = 6, LD32_2PASS = 7 };

template <int KIND>
__global__ void __launch_bounds__(512) k(long long* out, float* sink, int iters) {
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) tmem_alloc<512>(smem_u32(&slot));
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(0.001f * (threadIdx.x + i));
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (KIND == LD32) {  // 4 x (32 lanes x 32 columns x 4 B) = 16 KB per warp per iteration
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tmem_ld32(lane_addr + ((it + c) & 3) * 32, r);
        tmem_ld_wait32(r);
        acc += __uint_as_float(r[c]);
      }
    } else if (KIND == LD32_2PASS) {  // issue two loads back to back, one wait
      uint32_t r2[32];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        tmem_ld32(lane_addr + ((it + c) & 1) * 64, r);
        tmem_ld32(lane_addr + ((it + c) & 1) * 64 + 32, r2);
        tmem_ld_wait32(r);
        acc += __uint_as_float(r[c]) + __uint_as_float(r2[c]);
      }
    } else if (KIND == ST16) {  // 4 x (32 lanes x 16 columns x 4 B) = 8 KB per warp per iteration
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_st16(lane_addr + c * 16, r);
      tmem_st_wait();
    } else if (KIND == EX2) {  // 32 independent ex2 per thread per iteration
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float e;
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(__uint_as_float(r[i])));
        r[i] = __float_as_uint(e);
      }
    } else if (KIND == FFMA2) {  // 16 packed FMAs per thread per iteration (32 scalar FMAs)
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        uint64_t a, b, d;
        asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "r"(r[i]), "r"(r[i + 1]));
        asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(1.0001f), "f"(1.0001f));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %1;" : "=l"(d) : "l"(a), "l"(b));
        asm("mov.b64 {%0, %1}, %2;" : "=r"(r[i]), "=r"(r[i + 1]) : "l"(d));
      }
    } else if (KIND == CVT) {  // 16 cvt.f16x2 per iteration
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        uint32_t o;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(o) : "f"(__uint_as_float(r[i + 1])), "f"(__uint_as_float(r[i])));
        r[i] ^= o;
      }
    } else if (KIND == MAX3) {  // 16 three-input max per iteration, 4 chains
      float m0 = acc, m1 = acc, m2 = acc, m3 = acc;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(m0) : "f"(__uint_as_float(r[i])), "f"(__uint_as_float(r[i + 1])));
        asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(m1) : "f"(__uint_as_float(r[i + 2])), "f"(__uint_as_float(r[i + 3])));
        asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(m2) : "f"(__uint_as_float(r[i + 4])), "f"(__uint_as_float(r[i + 5])));
        asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(m3) : "f"(__uint_as_float(r[i + 6])), "f"(__uint_as_float(r[i + 7])));
      }
      acc = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
    } else if (KIND == MIX) {  // the exp pass of one 32-column chunk: ld32, 16 FFMA2, 32 ex2, 16 cvt, st16
      tmem_ld32(lane_addr + (it & 3) * 32, r);
      tmem_ld_wait32(r);
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        uint64_t a, b, d;
        asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "r"(r[i]), "r"(r[i + 1]));
        asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(0.5f), "f"(0.5f));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %2;" : "=l"(d) : "l"(a), "l"(b));
        float e0, e1;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(e0), "=f"(e1) : "l"(d));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(e0));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(e1));
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(pk[i / 2]) : "f"(e1), "f"(e0));
      }
      tmem_st16(lane_addr + 256 + (it & 3) * 16, pk);
      tmem_st_wait();
    }
  }
  const long long t1 = clock64();
#pragma unroll
  for (int i = 0; i < 32; ++i) acc += __uint_as_float(r[i]);
  if (acc == 12345.678f) sink[0] = acc;
  if (threadIdx.x == 0) out[0] = t1 - t0;
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}

template <int KIND>
void run(const char* name, double unit_per_warp_iter, const char* unit) {
  long long* out; float* sink;
  cudaMalloc(&out, 8); cudaMalloc(&sink, 4);
  const int iters = 2000;
  for (int warps : {4, 8, 16}) {
    k<KIND><<<1, warps * 32>>>(out, sink, iters);
    k<KIND><<<1, warps * 32>>>(out, sink, iters);
    long long cyc = 0;
    cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    printf("%-12s %2d warps: %8.1f cycles/iter  -> %7.1f %s per clk per SM\n", name, warps, (double)cyc / iters,
           unit_per_warp_iter * warps * iters / (double)cyc, unit);
  }
}

int main() {
  run<LD32>("tmem ld x32", 4 * 4096.0, "B");
  run<LD32_2PASS>("tmem ld 2x", 4 * 4096.0, "B");
  run<ST16>("tmem st x16", 4 * 2048.0, "B");
  run<EX2>("mufu ex2", 32 * 32.0, "ex2");
  run<FFMA2>("ffma2", 16 * 32.0, "ffma2 lanes");
  run<CVT>("cvt f16x2", 16 * 32.0, "cvt lanes");
  run<MAX3>("fmnmx3", 16 * 32.0, "max3 lanes");
  run<MIX>("exp chunk", 32 * 32.0, "elements");
  return 0;
}
