// Microbenchmark (diagnostic): cost of issuing tcgen05.mma from straight-line code with precomputed
// descriptors (one IADD per operand per MMA).
#include <cstdio>
#include "sm100.cuh"
using namespace tb;

__device__ __forceinline__ uint64_t desc_pack(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}

template <int NMMA, int N, bool TS, int AMN = 0>
__global__ void __launch_bounds__(128) k(long long* out) {
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
  if (threadIdx.x < 32) {
    constexpr uint32_t idesc = umma_idesc_f16(128, N, AMN, (TS || AMN) ? 1 : 0);
    const uint32_t a = smem_u32(smem), b = smem_u32(smem + 32768);
    const uint64_t da = umma_desc_sw128(a, AMN ? 16384 : 16, 1024), db = umma_desc_sw128(b, (TS || AMN) ? 16384 : 16, 1024);
    const uint32_t alo = (uint32_t)da, blo = (uint32_t)db, hi = (uint32_t)(da >> 32);
    for (int rep = 0; rep < 3; ++rep) {
      long long t0 = clock64();
      if (elect_one()) {
#pragma unroll
        for (int i = 0; i < NMMA; ++i) {
          if (TS) umma_f16_ts(tmem + 256, tmem + (i & 7) * 8, desc_pack(blo + (i & 7) * 128, hi), idesc, i > 0);
          else if (AMN) umma_f16_ss(tmem + 256, desc_pack(alo + (i & 7) * 128, hi), desc_pack(blo + (i & 7) * 128, hi), idesc, i > 0);
          else umma_f16_ss(tmem + 256, desc_pack(alo + (i & 3) * 2, hi), desc_pack(blo + (i & 3) * 2, hi), idesc, i > 0);
        }
      }
      __syncwarp();
      long long t1 = clock64();
      if (elect_one()) umma_commit(smem_u32(&bar));
      __syncwarp();
      mbar_wait(smem_u32(&bar), rep & 1);
      long long t2 = clock64();
      if (rep == 2 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}

template <int NMMA, int N, bool TS, int AMN = 0>
void run(long long* d) {
  cudaFuncSetAttribute(k<NMMA, N, TS, AMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  k<NMMA, N, TS, AMN><<<1, 128, 70000>>>(d);
  if (AMN) printf("[A MN-major, B MN-major] ");
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  printf("%s N=%3d n_mma=%2d: issue %5lld cyc (%.1f/mma), retire %5lld cyc (%.1f/mma) %s\n", TS ? "TS" : "SS", N, NMMA, h[0],
         (double)h[0] / NMMA, h[1], (double)h[1] / NMMA, e ? cudaGetErrorString(e) : "");
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  run<1, 48, false>(d); run<8, 48, false>(d); run<16, 48, false>(d); run<32, 48, false>(d);
  run<8, 16, false>(d); run<16, 16, false>(d); run<16, 128, false>(d); run<32, 128, false>(d); run<16, 256, false>(d);
  run<8, 48, true>(d); run<16, 48, true>(d); run<32, 48, true>(d); run<16, 64, true>(d); run<16, 16, true>(d);
  run<32, 16, true>(d); run<32, 32, true>(d); run<32, 64, true>(d); run<32, 80, true>(d); run<32, 96, true>(d); run<32, 128, true>(d);
  run<32, 64, false>(d); run<32, 32, false>(d);
  run<32, 48, false, 1>(d); run<32, 64, false, 1>(d); run<32, 128, false, 1>(d);
  return 0;
}
