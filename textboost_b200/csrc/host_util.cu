#include "host_util.h"

#include <stdarg.h>

namespace tb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return TB_E_CUDA;
  }
  return TB_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
  }
  return fn;
}

int make_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box) {
  EncodeTiledFn enc = get_encode();
  TB_REQUIRE(enc, TB_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
#ifdef TB_BF16
  const CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
#else
  const CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
#endif
  CUresult r = enc(out, dt, (cuuint32_t)rank, const_cast<void*>(base),
                   gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank=%d dims=[%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u]",
              (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
              box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return TB_E_CUDA;
  }
  return TB_OK;
}

int make_tmap_f32_plain(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                        const uint64_t* strides_bytes, const uint32_t* box) {
  EncodeTiledFn enc = get_encode();
  TB_REQUIRE(enc, TB_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx,
                   es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(f32) failed (%d): rank=%d dims=[%llu,%llu,%llu] box=[%u,%u,%u]", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0);
    return TB_E_CUDA;
  }
  return TB_OK;
}

constexpr int MAX_WS = 64;  // PyTorch hands out streams from a pool of 32 per device and priority
static Workspace g_ws[MAX_WS];
static int g_nws = 0;

static int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) cudaGetLastError();
  return dev;
}

const Workspace* find_ws(void* stream) {
  const int dev = current_device();
  for (int i = 0; i < g_nws; ++i)
    if (g_ws[i].stream == stream && g_ws[i].device == dev) return &g_ws[i];
  return nullptr;
}

}  // namespace tb

using namespace tb;

extern "C" int tb_set_workspace(void* stream, void* ptr, size_t bytes) {
  int rc = tb_check_device();
  if (rc) return rc;
  TB_REQUIRE(ptr == nullptr || ((uintptr_t)ptr % 256 == 0 && bytes > 2 * WS_COUNTER_BYTES), TB_E_ARG,
             "tb_set_workspace: pointer must be 256-byte aligned and larger than %zu bytes",
             2 * WS_COUNTER_BYTES);
  int slot = -1;
  const int dev = current_device();
  for (int i = 0; i < g_nws; ++i)
    if (g_ws[i].stream == stream && g_ws[i].device == dev) slot = i;
  if (slot < 0) {
    TB_REQUIRE(ptr != nullptr, TB_E_ARG, "tb_set_workspace: no workspace registered for this stream");
    TB_REQUIRE(g_nws < MAX_WS, TB_E_ARG, "tb_set_workspace: more than %d streams", MAX_WS);
    slot = g_nws++;
  }
  g_ws[slot].device = dev;
  g_ws[slot].stream = stream;
  g_ws[slot].base = (char*)ptr;
  g_ws[slot].bytes = ptr ? bytes : 0;
  if (ptr) {
    cudaError_t e = cudaMemsetAsync(ptr, 0, WS_COUNTER_BYTES, (cudaStream_t)stream);
    TB_REQUIRE(e == cudaSuccess, TB_E_CUDA, "tb_set_workspace memset: %s", cudaGetErrorString(e));
  } else {
    g_ws[slot] = g_ws[--g_nws];
  }
  return TB_OK;
}


extern "C" int tb_version(void) { return TB_ABI_VERSION; }

extern "C" int tb_storage_dtype(void) {
#ifdef TB_BF16
  return TB_STORAGE_BF16;
#else
  return TB_STORAGE_F16;
#endif
}

extern "C" const char* tb_last_error(void) { return tb::g_err; }

extern "C" int tb_check_device(void) {
  static int cached = 1;  // 1 = unknown
  if (cached != 1) {
    if (cached != 0) tb::set_error("textboost_b200 requires an sm_100 (B200) device; no fallback path exists");
    return cached;
  }
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    tb::set_error("no CUDA device: textboost_b200 has no CPU fallback");
    cudaGetLastError();
    return TB_E_ARCH;  // not cached: a device may appear later in another process state
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    tb::set_error("device compute capability %d.%d is not sm_100; textboost_b200 has no fallback path",
                  major, minor);
    cached = TB_E_ARCH;
    return cached;
  }
  cached = 0;
  return 0;
}
