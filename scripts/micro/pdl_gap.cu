// Microbenchmark (diagnostic): per-launch gap of a chain of dependent kernels inside a CUDA graph, with plain
// stream-order edges vs programmatic dependent launch (griddepcontrol.wait at the top of the consumer,
// griddepcontrol.launch_dependents early in the producer).  Kernels: a ~WORK-cycle body on `ctas` CTAs.
#include <cstdio>
#include <cuda_runtime.h>

template <bool PDL>
__global__ void k(float* buf, int work) {
  if (PDL) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  __shared__ float s[1024];
  s[threadIdx.x] = threadIdx.x;  // "prologue" independent of the previous kernel
  __syncthreads();
  if (PDL) asm volatile("griddepcontrol.wait;" ::: "memory");
  float v = buf[blockIdx.x * blockDim.x + threadIdx.x];
  const long long t0 = clock64();
  while (clock64() - t0 < work) v = v * 1.0001f + s[(threadIdx.x + 1) & 1023] * 1e-9f;
  buf[blockIdx.x * blockDim.x + threadIdx.x] = v;
}

template <bool PDL>
float run(float* buf, int n_kernels, int ctas, int work, cudaStream_t st) {
  cudaGraph_t g;
  cudaGraphExec_t ge;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < n_kernels; ++i) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = PDL ? 1 : 0;
    cudaLaunchKernelEx(&cfg, k<PDL>, buf, work);
  }
  cudaStreamEndCapture(st, &g);
  cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
  if (e != cudaSuccess) { printf("instantiate: %s\n", cudaGetErrorString(e)); return -1; }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) cudaGraphLaunch(ge, st);
  cudaEventRecord(e0, st);
  for (int i = 0; i < 10; ++i) cudaGraphLaunch(ge, st);
  cudaEventRecord(e1, st);
  cudaStreamSynchronize(st);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  e = cudaGetLastError();
  if (e != cudaSuccess) printf("error: %s\n", cudaGetErrorString(e));
  return ms / 10 / n_kernels * 1e3f;  // us per kernel
}

int main() {
  float* buf;
  cudaMalloc(&buf, 148 * 8 * 256 * 4);
  cudaMemset(buf, 0, 148 * 8 * 256 * 4);
  cudaStream_t st;
  cudaStreamCreate(&st);
  for (int ctas : {148, 592}) {
    for (int work : {2000, 10000, 40000}) {
      float a = run<false>(buf, 200, ctas, work, st), b = run<true>(buf, 200, ctas, work, st);
      printf("ctas %4d body %6d cycles (%.1f us @1.9GHz): plain %.2f us/kernel, PDL %.2f us/kernel, saved %.2f us\n", ctas,
             work, work / 1900.0, a, b, a - b);
    }
  }
  return 0;
}
