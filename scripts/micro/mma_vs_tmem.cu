// Microbenchmark (diagnostic): does tcgen05.ld / tcgen05.st / MUFU traffic from other warps of the CTA slow tcgen05.mma down?
// warp 8 issues NMMA MMAs (SS N=128 "S = Q K^T" or TS N=48 "O += P V"); warps 0..7 run a background loop:
// 0 = idle, 1 = tcgen05.ld x32 in a loop, 2 = tcgen05.st x16 loop, 3 = MUFU loop, 4 = ld + mufu + st (softmax-like).
#include <cstdio>
#include "sm100.cuh"
using namespace tb;

__device__ __forceinline__ uint64_t desc_pack(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}

template <int NMMA, int N, bool TS, int BG>
__global__ void __launch_bounds__(288) k(long long* out, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  for (int i = threadIdx.x; i < 65536 / 4; i += 288) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); stop = 0; }
  fence_async_smem();
  if (threadIdx.x < 32) tmem_alloc<512>(smem_u32(&slot));
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 8) {
    constexpr uint32_t idesc = umma_idesc_f16(128, N, 0, TS ? 1 : 0);
    const uint32_t a = smem_u32(smem), b = smem_u32(smem + 32768);
    const uint64_t da = umma_desc_sw128(a, 16, 1024), db = umma_desc_sw128(b, TS ? 16384 : 16, 1024);
    const uint32_t alo = (uint32_t)da, blo = (uint32_t)db, hi = (uint32_t)(da >> 32);
    for (int rep = 0; rep < 3; ++rep) {
      long long t0 = clock64();
      if (elect_one()) {
#pragma unroll
        for (int i = 0; i < NMMA; ++i) {
          if (TS) umma_f16_ts(tmem + 384, tmem + 256 + (i & 7) * 8, desc_pack(blo + (i & 7) * 128, hi), idesc, i > 0);
          else umma_f16_ss(tmem + 384, desc_pack(alo + (i & 3) * 2, hi), desc_pack(blo + (i & 3) * 2, hi), idesc, i > 0);
        }
      }
      __syncwarp();
      long long t1 = clock64();
      if (elect_one()) umma_commit(smem_u32(&bar));
      __syncwarp();
      mbar_wait(smem_u32(&bar), rep & 1);
      long long t2 = clock64();
      if (rep == 2 && (threadIdx.x & 31) == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    stop = 1;
  } else {
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(-0.001f * (threadIdx.x + i));
    float acc = 0.f;
    int it = 0;
    while (BG != 0 && !stop) {
      if (BG == 1 || BG == 4) {
        tmem_ld32(lane_addr + (it & 3) * 32, r);
        tmem_ld_wait32(r);
      }
      if (BG == 3 || BG == 4) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float e;
          asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(__uint_as_float(r[i])));
          r[i] = __float_as_uint(-e);
        }
      }
      if (BG == 2 || BG == 4) {
        tmem_st16(lane_addr + 128 + (it & 3) * 16, r);
        tmem_st_wait();
      }
      acc += __uint_as_float(r[it & 31]);
      ++it;
    }
    if (acc == 1234.5f) sink[0] = acc;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}

template <int NMMA, int N, bool TS, int BG>
void run(long long* d, float* sink) {
  cudaFuncSetAttribute(k<NMMA, N, TS, BG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  k<NMMA, N, TS, BG><<<1, 288, 70000>>>(d, sink);
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  const char* bg[] = {"idle", "tmem ld", "tmem st", "mufu", "ld+mufu+st"};
  printf("%s N=%3d n_mma=%2d background %-10s: issue %.1f/mma, retire %.1f/mma %s\n", TS ? "TS" : "SS", N, NMMA, bg[BG],
         (double)h[0] / NMMA, (double)h[1] / NMMA, e ? cudaGetErrorString(e) : "");
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  float* sink; cudaMalloc(&sink, 4);
  run<64, 128, false, 0>(d, sink); run<64, 128, false, 1>(d, sink); run<64, 128, false, 2>(d, sink);
  run<64, 128, false, 3>(d, sink); run<64, 128, false, 4>(d, sink);
  run<64, 48, true, 0>(d, sink); run<64, 48, true, 1>(d, sink); run<64, 48, true, 2>(d, sink);
  run<64, 48, true, 3>(d, sink); run<64, 48, true, 4>(d, sink);
  return 0;
}
