"""TEST INFRASTRUCTURE ONLY: compiles the KERNEL SOURCE of the byte-exact image kernels (csrc/image.cu, csrc/augment.cu)
for the host with g++ and runs it single-threaded, so that the CPU suite checks the very lines of CUDA C that run on the
GPU — index arithmetic, integer accumulation, rounding order — against PIL / torchvision, not a transcription of them.

How: the text of each .cu file up to the end of its `namespace tb { ... }` block (kernels and device helpers; the
extern "C" launch wrappers are left out) is compiled behind a shim that defines __global__ / __device__ away, provides
blockIdx / blockDim / gridDim / threadIdx for ONE thread of ONE block (every kernel here is a grid-stride loop, so that
thread visits every element), maps __ldg to a load and the explicitly rounded intrinsics (__dadd_rn, __fmul_rn, ...) to
the plain IEEE operation under -ffp-contract=off.  Nothing here is on the product path."""
import ctypes
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "textboost_b200", "csrc")

SHIM = r"""
#include <cmath>
#include <cstdint>
#include <cstring>
#define __global__
#define __device__
#define __forceinline__ inline
#define __launch_bounds__(...)
struct emu_dim3 { unsigned x, y, z; };
static emu_dim3 blockIdx = {0, 0, 0}, threadIdx = {0, 0, 0}, blockDim = {1, 1, 1}, gridDim = {1, 1, 1};
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
namespace tb { static inline int num_sms() { return 1; } }
"""

DRIVERS = r"""
extern "C" void emu_resample_h(const unsigned char* src, long long stride, const int* bounds, const int* kk, int ksize,
                               int row0, int nrows, int left, int cw, int channels, unsigned char* mid) {
  tb::resample_h_u8_kernel(src, stride, bounds, kk, ksize, row0, nrows, left, cw, channels, mid);
}
extern "C" void emu_resample_v(const unsigned char* mid, const int* bounds, const int* kk, int ksize, int row0, int top,
                               int ch, int cw, int channels, float scale, float mean, float std, float* out,
                               unsigned char* u8) {
  tb::resample_v_norm_kernel(mid, bounds, kk, ksize, row0, top, ch, cw, channels, scale, mean, std, out, u8);
}
extern "C" void emu_gather(const unsigned char* src, int sh, int sw, int channels, unsigned char* out, int oh, int ow,
                           int ox, int oy, int clamp, int flip, int tw, int th, int frame) {
  tb::img_gather_u8_kernel(src, sh, sw, channels, out, oh, ow, ox, oy, clamp, flip, tw, th, frame);
}
extern "C" void emu_affine(const unsigned char* src, int H, int W, int channels, unsigned char* out, const double* m,
                           int bicubic) {
  tb::img_affine_u8_kernel(src, H, W, channels, out, m[0], m[1], m[2], m[3], m[4], m[5], bicubic);
}
extern "C" void emu_grayscale(const unsigned char* src, unsigned char* out, long long npix) {
  tb::img_grayscale_u8_kernel(src, out, npix);
}
"""


# second host build: the fp16 elementwise kernels of csrc/vae.cu and csrc/sampler.cu (grid-stride ones only; the row
# softmax is a cooperative 256-thread kernel and is not run here).  __half is the compiler's IEEE binary16 (_Float16).
SHIM_F16 = r"""
typedef _Float16 __half;
struct __half2 { __half x, y; };
struct float2 { float x, y; };
struct uint4 { unsigned x, y, z, w; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline float __half2float(__half h) { return (float)h; }
static inline __half __float2half_rn(float f) { return (__half)f; }
static inline __half __float2half(float f) { return (__half)f; }
static inline float2 __half22float2(__half2 h) { return float2{(float)h.x, (float)h.y}; }
namespace tb {  // csrc/sm100.cuh: the 16-bit type of the build (this host build is the fp16 one) and its conversions
using half_t = __half;
using half2_t = __half2;
static inline float h2f(half_t h) { return (float)h; }
static inline half_t f2h(float f) { return (half_t)f; }
static inline float2 h22f2(half2_t h) { return float2{(float)h.x, (float)h.y}; }
static inline half2_t ff2h2(float a, float b) { return half2_t{(half_t)a, (half_t)b}; }
}
#define __shared__ static
#define __expf(x) expf(x)
static inline void __syncthreads() {}
static inline float __shfl_xor_sync(unsigned, float v, int) { return v; }
namespace tb {
static inline uint32_t pack_half2(float a, float b) {
  __half2 h{(__half)a, (__half)b};
  uint32_t u;
  memcpy(&u, &h, 4);
  return u;
}
}
"""

DRIVERS_F16 = r"""
extern "C" void emu_im2col_pad(const void* x, void* col, int B, int H, int W, int C, int pad_lo) {
  tb::im2col3x3s2_pad_kernel((const __half*)x, (__half*)col, B, H, W, C, pad_lo);
}
extern "C" void emu_vae_sample(const void* m, long long ld, const float* eps, float* lat, float* mean, float* sd, int HW,
                               int L, long long total, float scale) {
  tb::vae_sample_kernel((const __half*)m, ld, eps, lat, mean, sd, HW, L, total, scale);
}
extern "C" void emu_dpm_cfg_step(float* x, const void* eps, const float* m_prev, float* m_out, void* unet_in, long long n,
                                 float g, float a_i, float s_i, int v_pred, float c_x, float c_d0, float c_d1) {
  tb::dpm_cfg_step_kernel(x, (const __half*)eps, m_prev, m_out, (__half*)unet_in, n, g, a_i, s_i, v_pred, c_x, c_d0, c_d1);
}
extern "C" void emu_vae_decode_in(const float* lat, const float* w, const float* b, void* z, int L, int HW,
                                  long long total, float inv_scale) {
  tb::vae_decode_in_kernel(lat, w, b, (__half*)z, L, HW, total, inv_scale);
}
extern "C" void emu_image_u8(const void* x, long long ld, unsigned char* out, long long npix, int channels) {
  tb::image_u8_kernel((const __half*)x, ld, out, npix, channels);
}
"""

# third host build: the grid-stride kernels of csrc/elementwise.cu (core path: GEGLU, nearest upsample and its backward,
# strided copy / accumulate, fp32->fp16 cast, the stride-2 window gather and zero stuffing, sinusoidal timestep
# embedding, SiLU, add_noise / velocity target, the direct 4-channel convolutions).  `extern __shared__` staging arrays
# become one global array; warp-cooperative kernels (mse reduction, conv_out) compile but are not run.
SHIM_ELEMENTWISE = SHIM_F16.replace("#define __shared__ static", "#define __shared__") + r"""
struct float4 { float x, y, z, w; };
static inline float __fdividef(float a, float b) { return a / b; }
static inline float atomicAdd(float* p, float v) { float o = *p; *p += v; return o; }
namespace tb { __half sw[1 << 17]; }
"""

DRIVERS_ELEMENTWISE = r"""
#define H16(p) ((__half*)(p))
#define CH16(p) ((const __half*)(p))
extern "C" void emu_geglu_fwd(const void* h, void* out, long long M, int F) { tb::geglu_fwd_kernel(CH16(h), H16(out), M, F); }
extern "C" void emu_geglu_bwd(const void* dg, const void* h, void* dh, long long M, int F) {
  tb::geglu_bwd_kernel(CH16(dg), CH16(h), H16(dh), M, F);
}
extern "C" void emu_upsample_fwd(const void* x, void* y, int B, int H, int W, int C) {
  tb::upsample2x_fwd_kernel(CH16(x), H16(y), B, H, W, C);
}
extern "C" void emu_upsample_bwd(const void* dy, void* dx, int B, int H, int W, int C) {
  tb::upsample2x_bwd_kernel(CH16(dy), H16(dx), B, H, W, C);
}
extern "C" void emu_copy2d(void* dst, long long ldd, const void* src, long long lds, long long rows, int cols, int acc) {
  tb::copy2d_kernel(H16(dst), ldd, CH16(src), lds, rows, cols, acc);
}
extern "C" void emu_cast(void* dst, long long ldd, const float* src, long long lds, long long rows, int cols, float scale) {
  tb::cast_f32_f16_kernel(H16(dst), ldd, src, lds, rows, cols, scale);
}
extern "C" void emu_im2col(const void* x, void* col, int B, int H, int W, int C) {
  tb::im2col3x3s2_kernel(CH16(x), H16(col), B, H, W, C);
}
extern "C" void emu_zero_stuff(const void* dy, void* out, int B, int Ho, int Wo, int C) {
  tb::zero_stuff2x_kernel(CH16(dy), H16(out), B, Ho, Wo, C);
}
extern "C" void emu_timestep_embedding(const long long* t, void* out, int B, int dim) {
  for (int i = 0; i < B * dim / 2; ++i) {  // not a grid-stride kernel: one emulated block per element
    blockIdx.x = i;
    tb::timestep_embedding_kernel(t, H16(out), B, dim);
  }
  blockIdx.x = 0;
}
extern "C" void emu_silu(const void* x, void* y, long long n8) { tb::silu_kernel(CH16(x), H16(y), n8); }
extern "C" void emu_add_noise(const float* x0, const float* eps, const long long* t, const float* acp, void* noisy,
                              float* target, int per_image, long long n, int v_pred) {
  tb::add_noise_kernel(x0, eps, t, acp, H16(noisy), target, per_image, n, v_pred);
}
extern "C" void emu_conv_in(const void* x, const void* w, const void* bias, void* y, int B, int H, int W, int Cin,
                            int Cout) {
  tb::conv_in_kernel(CH16(x), CH16(w), CH16(bias), H16(y), B, H, W, Cin, Cout);
}
extern "C" void emu_conv_out_bwd4(const void* dy, const void* w, void* dh, int B, int H, int W, int Cin) {
  tb::conv_out_bwd_kernel<4>(CH16(dy), CH16(w), H16(dh), B, H, W, Cin);
}
"""

_lib_elementwise = None


def lib_elementwise():
    global _lib_elementwise
    if _lib_elementwise is None:
        d = tempfile.mkdtemp(prefix="tb_kernel_emu_ew_")
        src = os.path.join(d, "emu_ew.cpp")
        with open(src, "w") as f:
            f.write(SHIM + SHIM_ELEMENTWISE + _kernel_text("elementwise.cu") + "\n" + DRIVERS_ELEMENTWISE)
        so = os.path.join(d, "emu_ew.so")
        subprocess.run(["g++", "-O1", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-std=c++17",
                        "-Wno-unknown-pragmas", "-o", so, src], check=True, capture_output=True, text=True)
        _lib_elementwise = ctypes.CDLL(so)
    return _lib_elementwise


# fourth host build: the AdamW arithmetic of csrc/optim.cu (mixing mask, inf / norm reduction, unscale + clip + AdamW
# update).  With one emulated lane per warp the shuffles contribute nothing and the vote is the lane's own predicate; the
# finishing kernel (one warp per embedding row, block-wide scalar bookkeeping) is cooperative and is not run here.
SHIM_OPTIM = r"""
static inline float __shfl_xor_sync(unsigned, float v, int) { (void)v; return 0.f; }
static inline int __any_sync(unsigned, int p) { return p; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline float atomicAdd(float* p, float v) { float o = *p; *p += v; return o; }
static inline void __syncthreads() {}
#define __shared__ static
using std::isfinite;
"""

DRIVERS_OPTIM = r"""
extern "C" void emu_mix_mask(float* g, long long n, int D, int r, int parity) { tb::optim_mix_mask_kernel(g, n, D, r, parity); }
extern "C" void emu_adamw_reduce_update(float* p, float* g, float* m, float* v, long long n_lora, int n_rows, int D,
                                        float lr_lora, float lr_emb, float beta1, float beta2, float eps, float wd,
                                        float max_norm, float inv_world, float* state) {
  tb::AdamCfg c;
  c.n_lora = n_lora; c.n_rows = n_rows; c.D = D; c.n_total = n_lora + (long long)n_rows * D;
  c.lr_lora = lr_lora; c.lr_emb = lr_emb; c.beta1 = beta1; c.beta2 = beta2; c.eps = eps; c.wd = wd;
  c.max_norm = max_norm; c.inv_world = inv_world; c.mean_norm = 0.f;
  c.growth_factor = 2.f; c.backoff_factor = 0.5f; c.growth_interval = 2000;
  c.sched_kind = 0; c.sched_warmup = 0.f; c.sched_total = 0.f;
  tb::optim_reduce_kernel(g, c, state);
  tb::optim_update_kernel(p, g, m, v, c, state);
}
"""

_lib_optim = None


def lib_optim():
    global _lib_optim
    if _lib_optim is None:
        d = tempfile.mkdtemp(prefix="tb_kernel_emu_opt_")
        src = os.path.join(d, "emu_opt.cpp")
        with open(src, "w") as f:
            f.write(SHIM + SHIM_OPTIM + _kernel_text("optim.cu") + "\n" + DRIVERS_OPTIM)
        so = os.path.join(d, "emu_opt.so")
        subprocess.run(["g++", "-O1", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-std=c++17",
                        "-Wno-unknown-pragmas", "-o", so, src], check=True, capture_output=True, text=True)
        _lib_optim = ctypes.CDLL(so)
    return _lib_optim


# fifth host build: the one-block-per-row / grid-stride kernels of csrc/clip.cu (token + position gather with the lazy
# decay scalar, sparse embedding-row gradient, LoRA weight packing, activation forward / backward, the TextBoostModel
# null-embedding override and its gradient mask).  Drivers walk blockIdx.x over the rows.
DRIVERS_CLIP = r"""
extern "C" void emu_clip_embed(const long long* ids, const float* base, const float* added, const float* decay,
                               const float* pos, float* x, int M, int L, int D, int n_base, int n_rows) {
  for (int m = 0; m < M; ++m) { blockIdx.x = m; tb::clip_embed_kernel(ids, base, added, decay, pos, x, M, L, D, n_base, n_rows); }
  blockIdx.x = 0;
}
extern "C" void emu_clip_embed_grad(const long long* ids, const float* g, float* rows, int M, int D, int n_base) {
  for (int m = 0; m < M; ++m) { blockIdx.x = m; tb::clip_embed_grad_kernel(ids, g, rows, M, D, n_base); }
  blockIdx.x = 0;
}
extern "C" void emu_lora_pack(const float* B, void* wext, void* wext_t, int nblk, int tmask, int D, int r, int rpad,
                              float scaling) {
  tb::lora_pack_kernel(B, (__half*)wext, (__half*)wext_t, nblk, tmask, D, r, rpad, scaling);
}
extern "C" void emu_act(const void* u, const void* g, void* out, long long n, int kind, int bwd) {
  if (bwd) tb::act_kernel<true>((const __half*)u, (const __half*)g, (__half*)out, n, kind);
  else tb::act_kernel<false>((const __half*)u, nullptr, (__half*)out, n, kind);
}
extern "C" void emu_null_override(const long long* ids, const float* null_emb, float* h, int B, int L, int D, int eos,
                                  int use_fixed, int bwd) {
  for (int m = 0; m < B * L; ++m) {
    blockIdx.x = m;
    if (bwd) tb::null_override_kernel<true>(ids, null_emb, h, L, D, eos, use_fixed);
    else tb::null_override_kernel<false>(ids, null_emb, h, L, D, eos, use_fixed);
  }
  blockIdx.x = 0;
}
"""

_lib_clip = None


def lib_clip():
    global _lib_clip
    if _lib_clip is None:
        d = tempfile.mkdtemp(prefix="tb_kernel_emu_clip_")
        src = os.path.join(d, "emu_clip.cpp")
        header = '#include "%s"\n' % os.path.join(ROOT, "include", "textboost_b200.h")
        with open(src, "w") as f:
            f.write(SHIM + SHIM_ELEMENTWISE + "#include <algorithm>\nusing std::min; using std::max;\n" + header
                    + _kernel_text("clip.cu") + "\n" + DRIVERS_CLIP)
        so = os.path.join(d, "emu_clip.so")
        subprocess.run(["g++", "-O1", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-std=c++17",
                        "-Wno-unknown-pragmas", "-o", so, src], check=True, capture_output=True, text=True)
        _lib_clip = ctypes.CDLL(so)
    return _lib_clip


# ---------------------------------------------------------------------------------------------------------------
# Block emulator for the COOPERATIVE SIMT kernels (shared memory, __syncthreads, warp shuffles, atomics): every CUDA
# thread of a block runs as an OS thread; __syncthreads is a block-wide std::barrier, shuffles exchange through a per-warp
# buffer between two per-warp barriers, a thread that returns drops out of both barriers, atomics are std::atomic_ref.
# Blocks run one after the other.  `__shared__` variables become function-level statics (one copy for the block),
# `extern __shared__` arrays a global buffer.  Slow and faithful: the tests use a few dozen small blocks.
SHIM_BLOCK = r"""
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>
#define __global__
#define __device__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __expf(x) expf(x)
using std::min; using std::max; using std::isfinite;
struct emu_dim3 { unsigned x, y, z; };
static thread_local emu_dim3 threadIdx = {0, 0, 0};
static emu_dim3 blockIdx = {0, 0, 0}, blockDim = {1, 1, 1}, gridDim = {1, 1, 1};
typedef _Float16 __half;
struct __half2 { __half x, y; };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct uint4 { unsigned x, y, z, w; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float __half2float(__half h) { return (float)h; }
static inline __half __float2half_rn(float f) { return (__half)f; }
static inline __half __float2half(float f) { return (__half)f; }
static inline float2 __half22float2(__half2 h) { return float2{(float)h.x, (float)h.y}; }
namespace tb {  // csrc/sm100.cuh: the 16-bit type of the build (this host build is the fp16 one) and its conversions
using half_t = __half;
using half2_t = __half2;
static inline float h2f(half_t h) { return (float)h; }
static inline half_t f2h(float f) { return (half_t)f; }
static inline float2 h22f2(half2_t h) { return float2{(float)h.x, (float)h.y}; }
static inline half2_t ff2h2(float a, float b) { return half2_t{(half_t)a, (half_t)b}; }
}
static inline float __fdividef(float a, float b) { return a / b; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline float atomicAdd(float* p, float v) { return std::atomic_ref<float>(*p).fetch_add(v); }
static std::barrier<>* emu_block_bar = nullptr;
static std::vector<std::unique_ptr<std::barrier<>>> emu_warp_bar;
static uint32_t emu_xchg[64][32];
static inline void __syncthreads() { emu_block_bar->arrive_and_wait(); }
template <typename T> static inline T emu_exchange(T v, int src_lane) {
  static_assert(sizeof(T) == 4, "32-bit shuffles only");
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  const int lanes = std::min(32u, blockDim.x - 32u * w);
  memcpy(&emu_xchg[w][l], &v, 4);
  emu_warp_bar[w]->arrive_and_wait();
  T r = v;
  if (src_lane >= 0 && src_lane < lanes) memcpy(&r, &emu_xchg[w][src_lane], 4);
  emu_warp_bar[w]->arrive_and_wait();
  return r;
}
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int o) { return emu_exchange(v, (int)(threadIdx.x & 31) ^ o); }
template <typename T> static inline T __shfl_sync(unsigned, T v, int src) { return emu_exchange(v, src); }
template <typename T> static inline T __shfl_down_sync(unsigned, T v, int d) { return emu_exchange(v, (int)(threadIdx.x & 31) + d); }
static inline int __any_sync(unsigned m, int p) {
  int r = 0;
  for (int l = 0; l < 32; ++l) r |= __shfl_sync(m, p, l);
  return r;
}
#define TB_REQUIRE(cond, code, ...) do { if (!(cond)) return (code); } while (0)
namespace tb {
static inline int num_sms() { return 148; }
static inline uint32_t pack_half2(float a, float b) {
  __half2 h{(__half)a, (__half)b};
  uint32_t u;
  memcpy(&u, &h, 4);
  return u;
}
float sg[1 << 16];
__half sw[1 << 17];
uint4 gsm[1 << 14];      // dynamic shared memory of gn_group_kernel
uint4 ln_smem[1 << 14];  // ... of ln_lora_fwd_kernel / ln_bwd_clip_kernel (the staged LoRA down-projection matrix)
}
// A process-wide pool of OS threads (never torn down: the workers sleep on a barrier until the process exits) runs every
// launch: worker t plays CUDA thread t of each block in turn; thread 0 installs the block's barriers between two
// launch-wide phases.  Blocks of more threads than the pool holds are not used by these kernels.
constexpr unsigned EMU_MAX_THREADS = 512;
struct EmuPool {
  const unsigned size;
  std::barrier<> go, done;
  const std::function<void(unsigned)>* job = nullptr;
  explicit EmuPool(unsigned n) : size(n), go(n + 1), done(n + 1) {
    for (unsigned t = 0; t < size; ++t)
      std::thread([this, t] {
        for (;;) {
          go.arrive_and_wait();
          (*job)(t);
          done.arrive_and_wait();
        }
      }).detach();
  }
  void run(const std::function<void(unsigned)>& j) {
    job = &j;
    go.arrive_and_wait();
    done.arrive_and_wait();
  }
};
// every launch wakes its whole pool twice, so the common blocks (<= 320 threads) get their own, smaller pool; the
// 512-thread pool exists only once a kernel needs it (the group-owner GroupNorm: 510 threads)
static EmuPool& emu_pool(unsigned threads) {
  static EmuPool* small = new EmuPool(320);
  static EmuPool* big = nullptr;
  if (threads <= 320) return *small;
  if (!big) big = new EmuPool(EMU_MAX_THREADS);
  return *big;
}
static void emu_launch(unsigned gx, unsigned gy, unsigned threads, const std::function<void()>& f) {
  if (threads > EMU_MAX_THREADS) abort();
  gridDim = {gx, gy, 1};
  blockDim = {threads, 1, 1};
  const unsigned warps = (threads + 31) / 32, nblocks = gx * gy;
  std::barrier<> phase(threads);
  std::unique_ptr<std::barrier<>> block_bar;
  emu_pool(threads).run([&](unsigned t) {
    if (t >= threads) return;
    threadIdx = {t, 0, 0};
    for (unsigned b = 0; b < nblocks; ++b) {
      if (t == 0) {
        blockIdx = {b % gx, b / gx, 0};
        block_bar.reset(new std::barrier<>(threads));
        emu_block_bar = block_bar.get();
        emu_warp_bar.clear();
        for (unsigned w = 0; w < warps; ++w)
          emu_warp_bar.emplace_back(new std::barrier<>(std::min(32u, threads - 32 * w)));
      }
      phase.arrive_and_wait();
      f();
      emu_warp_bar[t >> 5]->arrive_and_drop();  // a finished thread no longer takes part in the block's barriers
      emu_block_bar->arrive_and_drop();
      phase.arrive_and_wait();
    }
  });
}
// kernels without shared memory, barriers or shuffles: every CUDA thread is independent, so one host thread plays them
// all in turn (atomics stay atomic)
static void emu_launch_serial(unsigned gx, unsigned gy, unsigned threads, const std::function<void()>& f) {
  gridDim = {gx, gy, 1};
  blockDim = {threads, 1, 1};
  for (unsigned by = 0; by < gy; ++by)
    for (unsigned bx = 0; bx < gx; ++bx) {
      blockIdx = {bx, by, 0};
      for (unsigned t = 0; t < threads; ++t) {
        threadIdx = {t, 0, 0};
        f();
      }
    }
  threadIdx = {0, 0, 0};
}
"""

DRIVERS_NORM = r"""
extern "C" int emu_groupnorm_fwd(const void* x, const void* gamma, const void* beta, void* y, float* stats, int B, int HW,
                                 int C, int G, float eps, int silu) {
  tb::GnGeom g;
  if (tb::gn_geom(g, B, HW, C, G)) return -1;  // the launch geometry tb_groupnorm_fwd_f16 uses on a 148-SM part
  memset(stats, 0, sizeof(float) * B * G * 2);
  const unsigned gx = (HW + g.ppc - 1) / g.ppc, threads = g.nvec * g.k;
  emu_launch(gx, B, threads, [&] { tb::gn_stats_kernel<0, 4, 3>((const __half*)x, nullptr, nullptr, nullptr, nullptr, stats, g, eps, silu); });
  emu_launch(gx, B, threads, [&] { tb::gn_apply_kernel<0, 4, 3>((const __half*)x, nullptr, (const __half*)gamma, (const __half*)beta,
                                                          stats, nullptr, nullptr, (__half*)y, g, eps, silu); });
  return (int)threads;
}
extern "C" int emu_groupnorm_bwd(const void* dy, const void* x, const void* gamma, const void* beta, const float* stats,
                                 float* dstats, const void* add, void* dx, int B, int HW, int C, int G, float eps, int silu) {
  tb::GnGeom g;
  if (tb::gn_geom(g, B, HW, C, G)) return -1;
  memset(dstats, 0, sizeof(float) * B * G * 2);
  const unsigned gx = (HW + g.ppc - 1) / g.ppc, threads = g.nvec * g.k;
  emu_launch(gx, B, threads, [&] { tb::gn_stats_kernel<1, 2, 2>((const __half*)x, (const __half*)dy, (const __half*)gamma,
                                                          (const __half*)beta, stats, dstats, g, eps, silu); });
  emu_launch(gx, B, threads, [&] { tb::gn_apply_kernel<1, 2, 2>((const __half*)x, (const __half*)dy, (const __half*)gamma,
                                                          (const __half*)beta, stats, dstats, (const __half*)add, (__half*)dx,
                                                          g, eps, silu); });
  return (int)threads;
}
"""

_lib_norm = None


def _build_block(name, files, drivers, extra=""):
    d = tempfile.mkdtemp(prefix=f"tb_kernel_emu_{name}_")
    src = os.path.join(d, f"emu_{name}.cpp")
    header = '#include "%s"\n' % os.path.join(ROOT, "include", "textboost_b200.h")
    with open(src, "w") as f:
        kernels = "\n".join(_kernel_text(x) for x in files).replace("extern __shared__", "extern")
        f.write(SHIM_BLOCK + header + extra + kernels + "\n" + drivers)
    so = os.path.join(d, f"emu_{name}.so")
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-std=c++20", "-pthread",
                    "-Wno-unknown-pragmas", "-o", so, src], check=True, capture_output=True, text=True)
    return ctypes.CDLL(so)


def lib_norm():
    global _lib_norm
    if _lib_norm is None:
        _lib_norm = _build_block("norm", ["norm.cu"], DRIVERS_NORM)
    return _lib_norm


# ---------------------------------------------------------------------------------------------------------------
# The C ABI itself on the CPU: the WHOLE text of the SIMT source files (kernels AND their extern "C" entry points with
# the argument checks, launch geometry and dispatch) compiled behind the block emulator.  `kernel<<<grid, block, smem,
# stream>>>(args);` is rewritten textually into `emu_launch(grid, block, [&] { kernel(args); });`; the handful of runtime
# calls the entry points make (cudaMemsetAsync, cudaFuncSetAttribute, tb_check_device, check_launch, set_error) are
# shimmed.  The tcgen05 files (gemm.cu, attn.cu) are not part of it.
_COOPERATIVE_TOKENS = ("__syncthreads", "__shfl", "__shared__", "__any_sync", "warp_sum", "group_sum")


def _is_cooperative(kernel: str, text: str) -> bool:
    """Does the kernel's body use shared memory, barriers or warp shuffles (directly or through the warp-sum helpers)?"""
    import re
    base = kernel.split("<")[0].split("::")[-1]
    m = re.search(r"__global__[^;{]*?\b" + re.escape(base) + r"\s*\(", text)
    if not m:
        return True
    k = text.index("{", m.end())
    depth, q = 0, k
    while True:
        depth += {"{": 1, "}": -1}.get(text[q], 0)
        if depth == 0:
            break
        q += 1
    body = text[k:q]
    return any(tok in body for tok in _COOPERATIVE_TOKENS)


def _rewrite_launches(text: str) -> str:
    out, i = [], 0
    while True:
        j = text.find("<<<", i)
        if j < 0:
            out.append(text[i:])
            return "".join(out)
        k = j
        if text[k - 1] == ">":  # template argument list of the kernel
            depth = 0
            while True:
                k -= 1
                depth += {">": 1, "<": -1}.get(text[k], 0)
                if depth == 0:
                    break
        while text[k - 1].isalnum() or text[k - 1] in "_:":
            k -= 1
        name = text[k:j]
        e = text.index(">>>", j)
        parts, depth, cur = [], 0, ""
        for ch in text[j + 3:e]:  # launch configuration, split at top-level commas
            depth += {"(": 1, ")": -1}.get(ch, 0)
            if ch == "," and depth == 0:
                parts.append(cur)
                cur = ""
            else:
                cur += ch
        parts.append(cur)
        p = text.index("(", e)
        depth, q = 0, p
        while True:
            depth += {"(": 1, ")": -1}.get(text[q], 0)
            if depth == 0:
                break
            q += 1
        args = text[p + 1:q]
        out.append(text[i:k])
        launcher = "emu_launch" if _is_cooperative(name, text) else "emu_launch_serial"
        out.append(f"{launcher}({parts[0].strip()}, {parts[1].strip()}, [&] {{ {name}({args}); }})")
        i = q + 1


SHIM_ABI = r"""
#include <cstdarg>
#include <cstdio>
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
static char emu_last_error[512] = "";
extern "C" int tb_check_device(void) { return 0; }
extern "C" const char* emu_last_error_text(void) { return emu_last_error; }
namespace tb {
static inline void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(emu_last_error, sizeof(emu_last_error), fmt, ap);
  va_end(ap);
}
static inline int check_launch(const char*) { return 0; }
float smf[1 << 18];
}
#undef TB_REQUIRE
#define TB_REQUIRE(cond, code, ...) do { if (!(cond)) { tb::set_error(__VA_ARGS__); return (code); } } while (0)
static void emu_launch(dim3 grid, dim3 block, const std::function<void()>& f) { emu_launch(grid.x, grid.y, block.x, f); }
static void emu_launch_serial(dim3 grid, dim3 block, const std::function<void()>& f) {
  emu_launch_serial(grid.x, grid.y, block.x, f);
}
"""

ABI_FILES = ["norm.cu", "elementwise.cu", "clip.cu", "optim.cu", "vae.cu", "sampler.cu", "image.cu", "augment.cu",
             "unet_lora.cu"]
_lib_abi = {}


def lib_abi(bf16: bool = False):
    """The SIMT entry points of include/textboost_b200.h, built for the host from the product source (see above).
    bf16=True: the host twin of libtextboost_b200_bf16.so -- the 16-bit type of the build is the compiler's bfloat16
    (`__bf16`, round-to-nearest-even conversions) instead of binary16, everything else identical."""
    if bf16 not in _lib_abi:
        d = tempfile.mkdtemp(prefix="tb_abi_emu_")
        src = os.path.join(d, "abi.cpp")
        header = '#include "%s"\n' % os.path.join(ROOT, "include", "textboost_b200.h")
        body = []
        for name in ABI_FILES:
            text = open(os.path.join(CSRC, name)).read()
            text = "\n".join(l for l in text.splitlines() if not l.startswith("#include"))
            text = text.replace("extern __shared__", "extern").replace("#define TB_ENTER", "#undef TB_ENTER\n#define TB_ENTER")
            body.append(f"// ===== {name}\n" + _rewrite_launches(text))
        shim = SHIM_BLOCK
        if bf16:
            assert "typedef _Float16 __half;" in shim
            shim = shim.replace("typedef _Float16 __half;", "typedef __bf16 __half;  // the -DTB_BF16 build")
        with open(src, "w") as f:
            f.write(shim + header + SHIM_ABI + "\n".join(body))
        so = os.path.join(d, "abi.so")
        r = subprocess.run(["g++", "-O1", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-std=c++20",
                            "-pthread", "-Wno-unknown-pragmas", "-o", so, src], capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError(r.stderr[:6000])
        _lib_abi[bf16] = ctypes.CDLL(so)
    return _lib_abi[bf16]


_lib_f16 = None


def lib_f16():
    global _lib_f16
    if _lib_f16 is None:
        d = tempfile.mkdtemp(prefix="tb_kernel_emu16_")
        src = os.path.join(d, "emu16.cpp")
        with open(src, "w") as f:
            f.write(SHIM + SHIM_F16 + _kernel_text("vae.cu") + "\n" + _kernel_text("sampler.cu") + "\n" + DRIVERS_F16)
        so = os.path.join(d, "emu16.so")
        subprocess.run(["g++", "-O1", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-std=c++17",
                        "-Wno-unknown-pragmas", "-o", so, src], check=True, capture_output=True, text=True)
        _lib_f16 = ctypes.CDLL(so)
    return _lib_f16


def _kernel_text(name):
    text = open(os.path.join(CSRC, name)).read()
    end = text.index("}  // namespace tb") + len("}  // namespace tb")
    return "\n".join(l for l in text[:end].splitlines() if not l.startswith("#include"))


_lib = None


def lib():
    """ctypes handle of the host build of the image kernels (compiled once per process into a temp dir)."""
    global _lib
    if _lib is None:
        d = tempfile.mkdtemp(prefix="tb_kernel_emu_")
        src = os.path.join(d, "emu.cpp")
        with open(src, "w") as f:
            f.write(SHIM + _kernel_text("image.cu") + "\n" + _kernel_text("augment.cu") + "\n" + DRIVERS)
        so = os.path.join(d, "emu.so")
        subprocess.run(["g++", "-O1", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-std=c++17",
                        "-Wno-unknown-pragmas", "-o", so, src], check=True, capture_output=True, text=True)
        _lib = ctypes.CDLL(so)
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def fakes():
    """Drop-in replacements for tests/cabi_standin.py's numpy fakes that run the real kernel source instead."""
    import cabi_standin as S
    L = lib()

    def tb_resize_crop_normalize_u8(src, H, W, C, bx, kx, ksx, out_w, by, ky, ksy, out_h, row0, nrows, top, left, ch,
                                    cw, scale, mean, std, mid, out_f32, out_u8, stream):
        L.emu_resample_h(ctypes.c_void_p(S._addr(src)), ctypes.c_longlong(W * C), ctypes.c_void_p(S._addr(bx)),
                         ctypes.c_void_p(S._addr(kx)), ksx, row0, nrows, left, cw, C, ctypes.c_void_p(S._addr(mid)))
        L.emu_resample_v(ctypes.c_void_p(S._addr(mid)), ctypes.c_void_p(S._addr(by)), ctypes.c_void_p(S._addr(ky)),
                         ksy, row0, top, ch, cw, C, ctypes.c_float(scale), ctypes.c_float(mean), ctypes.c_float(std),
                         ctypes.c_void_p(S._addr(out_f32)), ctypes.c_void_p(S._addr(out_u8)))
        return 0

    def tb_img_gather_u8(src, sh, sw, C, out, oh, ow, ox, oy, clamp, flip, tw, th, frame, stream):
        L.emu_gather(ctypes.c_void_p(S._addr(src)), sh, sw, C, ctypes.c_void_p(S._addr(out)), oh, ow, ox, oy, clamp,
                     flip, tw, th, frame)
        return 0

    def tb_img_affine_u8(src, H, W, C, out, matrix6, bicubic, stream):
        L.emu_affine(ctypes.c_void_p(S._addr(src)), H, W, C, ctypes.c_void_p(S._addr(out)),
                     ctypes.c_void_p(S._addr(matrix6)), bicubic)
        return 0

    def tb_img_grayscale_u8(src, out, npix, stream):
        L.emu_grayscale(ctypes.c_void_p(S._addr(src)), ctypes.c_void_p(S._addr(out)), ctypes.c_longlong(npix))
        return 0

    return {f.__name__: f for f in (tb_resize_crop_normalize_u8, tb_img_gather_u8, tb_img_affine_u8,
                                    tb_img_grayscale_u8)}


def install(monkeypatch):
    import cabi_standin as S
    S.install(monkeypatch)
    monkeypatch.setattr(S, "FAKES", fakes())


def install_abi_bf16(monkeypatch):
    """install_abi with the bf16 host build and the process's precision policy switched to bf16 for the test (the
    ops layer then expects torch.bfloat16 tensors, as it does beside libtextboost_b200_bf16.so)."""
    from textboost_b200.precision import POLICY
    monkeypatch.setattr(POLICY, "name", "bf16")
    return install_abi(monkeypatch, bf16=True)


def install_abi(monkeypatch, bf16: bool = False):
    """Bind textboost_b200._cabi to the host build of the SIMT entry points for one test: `ops.*` / `C.call` then run the
    product's own ctypes signatures, argument checks, launch geometry and kernel source on CPU tensors.  The tcgen05
    entry points (GEMM, conv, attention) are absent and raise AttributeError."""
    from textboost_b200 import _cabi
    L = lib_abi(bf16)
    for name, argtypes in _cabi._SIGNATURES.items():
        if hasattr(L, name):
            fn = getattr(L, name)
            fn.argtypes = argtypes
            fn.restype = _cabi._RESTYPES.get(name, ctypes.c_int)
    L.emu_last_error_text.restype = ctypes.c_char_p
    L.tb_last_error = L.emu_last_error_text
    monkeypatch.setattr(_cabi, "_lib", L)
    monkeypatch.setattr(_cabi, "stream_ptr", lambda: ctypes.c_void_p(0))
    return L

