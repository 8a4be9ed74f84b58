// Microbenchmark (diagnostic, not product): issue->retire time of chains of tcgen05.mma on one SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I textboost_b200/csrc -o /tmp/mma_latency scripts/micro/mma_latency.cu
#include <cstdio>
#include "sm100.cuh"
using namespace tb;

// mode 0: SS dependent chain (same accumulator); 1: SS two alternating accumulators; 2: TS dependent; 3: TS alternating
__device__ __forceinline__ uint64_t desc_pack(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}

__global__ void __launch_bounds__(128) k(int N, int n_mma, int mode, int pre, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  fence_async_smem();
  if (threadIdx.x < 32) tmem_alloc<512>(smem_u32(&slot));
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_f16(128, N, 0, (mode >= 2) ? 1 : 0);
    const uint32_t a = smem_u32(smem), b = smem_u32(smem + 32768);
    const uint64_t da = umma_desc_sw128(a, 16, 1024), db = umma_desc_sw128(b, 16, 1024);
    const uint32_t alo = (uint32_t)da, blo = (uint32_t)db, hi = (uint32_t)(da >> 32);
    for (int rep = 0; rep < 3; ++rep) {
      long long t0 = clock64();
      for (int i = 0; i < n_mma; ++i) {
        const uint32_t d = tmem + 256 + ((mode & 1) ? (i & 1) * 128 : 0);
        if (mode < 2) {
          if (pre) umma_f16_ss(d, desc_pack(alo + (i & 3) * 2, hi), desc_pack(blo + (i & 3) * 2, hi), idesc, i > 1);
          else umma_f16_ss(d, umma_desc_sw128(a + (i & 3) * 32, 16, 1024), umma_desc_sw128(b + (i & 3) * 32, 16, 1024), idesc, i > 1);
        } else {
          if (pre) umma_f16_ts(d, tmem + (i & 7) * 8, desc_pack(blo + (i & 3) * 128, hi), idesc, i > 1);
          else umma_f16_ts(d, tmem + (i & 7) * 8, umma_desc_sw128(b + (i & 3) * 2048, 16384, 1024), idesc, i > 1);
        }
      }
      long long t1 = clock64();
      umma_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), rep & 1);
      long long t2 = clock64();
      if (rep == 2) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  const char* names[4] = {"SS dep", "SS alt", "TS dep", "TS alt"};
  for (int pre = 0; pre < 2; ++pre)
  for (int mode = 0; mode < 4; ++mode)
    for (int N : {16, 48, 128, 256})
      for (int n : {1, 8, 16}) {
        if ((mode & 1) && N > 128) continue;
        if (mode >= 2 && N > 128) continue;
        k<<<1, 128, 70000>>>(N, n, mode, pre, d);
        long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaGetLastError();
        printf("pre=%d %s N=%3d n_mma=%2d: issue %5lld cyc, retire %5lld cyc (%.0f cyc/mma)%s\n", pre, names[mode], N, n, h[0], h[1],
               (double)h[1] / n, e ? cudaGetErrorString(e) : "");
      }
  return 0;
}
