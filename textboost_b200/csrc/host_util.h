// Host-side helpers shared by the C-ABI entry points: thread-local error text,
// the driver entry point for cuTensorMapEncodeTiled, launch checking.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/textboost_b200.h"

namespace tb {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

// Tiled fp16 tensor map with 128-byte swizzle and zero OOB fill.
// dims/strides innermost first; strides in BYTES for dims 1..rank-1 (dim 0 is contiguous).
int make_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box);

// Tiled fp32 tensor map without swizzle (row-major box in shared memory): target of the TMA reduce-add that
// accumulates dQ in the attention backward.
int make_tmap_f32_plain(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                        const uint64_t* strides_bytes, const uint32_t* box);

// Per-stream scratch registered by the caller (tb_set_workspace): split-K partial tiles (gemm.cu) and the fp32
// dK/dV accumulators of the q-split attention backward (attn.cu).  The first WS_COUNTER_BYTES hold the split-K
// tile counters and stay zero between launches; everything after is free-for-all scratch, valid only between the
// start and end of one C-ABI call on that stream.
struct Workspace {
  int device;    // the table is keyed on (device, stream): the default stream's handle is 0 on every device
  void* stream;
  char* base;
  size_t bytes;
};
constexpr size_t WS_COUNTER_BYTES = 64 * 1024;
const Workspace* find_ws(void* stream);

inline int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

}  // namespace tb

#define TB_REQUIRE(cond, code, ...)  \
  do {                               \
    if (!(cond)) {                   \
      tb::set_error(__VA_ARGS__);    \
      return (code);                 \
    }                                \
  } while (0)
