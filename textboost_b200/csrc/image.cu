// Byte-exact image front end on the GPU: Pillow's antialiased 8-bit resize (the arithmetic of
// v2.Resize(size, LANCZOS) at /root/reference/textboost/dataset.py:326, 342), the crop (dataset.py:343-351) and the
// ToImage / ToDtype(scale) / Normalize(0.5, 0.5) chain (dataset.py:327-334), as two integer passes over a decoded uint8
// image that is resident in HBM.  The algorithm is Pillow's src/libImaging/Resample.c: per output index a window of
// fixed-point weights (22 fractional bits, computed on the host in double precision exactly as precompute_coeffs /
// normalize_coeffs_8bpc do and uploaded once per (input size, output size)), int32 accumulation from the rounding half,
// arithmetic shift, clip to [0, 255]; horizontal pass first, uint8 intermediate, then the vertical pass.  Only the
// rows / columns the crop window needs are produced.  HBM-bound byte work: a 1024^2 source is 3 MB, the passes read
// it once; no tensor cores, no shared-memory staging needed at these sizes (the weight tables are L1/L2-resident).
#include "host_util.h"

namespace tb {

constexpr int RESAMPLE_PRECISION_BITS = 32 - 8 - 2;

__device__ __forceinline__ int clip8(int v) {
  v >>= RESAMPLE_PRECISION_BITS;  // arithmetic shift, as Pillow's clip8 lookup indexes with a signed value
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// mid[r, j, c] = clip8(sum_x src[row0 + r, xmin(left + j) + x, c] * kk[left + j, x])   r < nrows, j < cw
__global__ void resample_h_u8_kernel(const unsigned char* __restrict__ src, long long src_stride,
                                     const int* __restrict__ bounds, const int* __restrict__ kk, int ksize, int row0,
                                     int nrows, int left, int cw, int channels, unsigned char* __restrict__ mid) {
  const long long total = (long long)nrows * cw * channels;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % channels);
    const long long t = i / channels;
    const int j = (int)(t % cw);
    const int r = (int)(t / cw);
    const int xx = left + j;
    const int xmin = __ldg(bounds + 2 * xx), n = __ldg(bounds + 2 * xx + 1);
    const int* k = kk + (long long)xx * ksize;
    const unsigned char* p = src + (long long)(row0 + r) * src_stride + (long long)xmin * channels + c;
    int ss = 1 << (RESAMPLE_PRECISION_BITS - 1);
    for (int x = 0; x < n; ++x) ss += (int)p[(long long)x * channels] * __ldg(k + x);
    mid[i] = (unsigned char)clip8(ss);
  }
}

// out[c, i, j] = ((float)clip8(sum_y mid[ymin(top + i) - row0 + y, j, c] * kk[top + i, y]) * scale - mean) / std
// in fp32 with separately rounded multiply, subtract and IEEE divide (no FMA contraction): the bits torchvision's
// ToDtype(float, scale=True) = x.float().mul_(1/255) followed by Normalize = sub_(mean).div_(std) produce; u8_out,
// when given, receives the resized + cropped bytes [ch, cw, channels] as well.
__global__ void resample_v_norm_kernel(const unsigned char* __restrict__ mid, const int* __restrict__ bounds,
                                       const int* __restrict__ kk, int ksize, int row0, int top, int ch, int cw,
                                       int channels, float scale, float mean, float std, float* __restrict__ out,
                                       unsigned char* __restrict__ u8_out) {
  const long long total = (long long)ch * cw * channels;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % channels);
    const long long t = i / channels;
    const int j = (int)(t % cw);
    const int r = (int)(t / cw);
    const int yy = top + r;
    const int ymin = __ldg(bounds + 2 * yy), n = __ldg(bounds + 2 * yy + 1);
    const int* k = kk + (long long)yy * ksize;
    const unsigned char* p = mid + ((long long)(ymin - row0) * cw + j) * channels + c;
    int ss = 1 << (RESAMPLE_PRECISION_BITS - 1);
    for (int y = 0; y < n; ++y) ss += (int)p[(long long)y * cw * channels] * __ldg(k + y);
    const int v = clip8(ss);
    if (u8_out) u8_out[i] = (unsigned char)v;
    if (out) out[((long long)c * ch + r) * cw + j] = __fdiv_rn(__fsub_rn(__fmul_rn((float)v, scale), mean), std);
  }
}

static inline unsigned image_grid_for(long long n) {
  long long b = (n + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace tb

using namespace tb;

extern "C" int tb_resize_crop_normalize_u8(const void* src_u8, int src_h, int src_w, int channels,
                                           const int32_t* bounds_x, const int32_t* kk_x, int ksize_x, int out_w,
                                           const int32_t* bounds_y, const int32_t* kk_y, int ksize_y, int out_h,
                                           int row0, int nrows, int top, int left, int crop_h, int crop_w,
                                           float scale, float mean, float std, void* mid_u8, float* out_f32_chw,
                                           void* out_u8_hwc,
                                           void* stream) {
  int rc = tb_check_device();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  TB_REQUIRE(src_u8 && bounds_x && kk_x && bounds_y && kk_y && mid_u8 && (out_f32_chw || out_u8_hwc), TB_E_ARG,
             "tb_resize_crop_normalize_u8: null pointer");
  TB_REQUIRE(channels >= 1 && channels <= 4 && src_h > 0 && src_w > 0 && out_w > 0 && out_h > 0, TB_E_SHAPE,
             "tb_resize_crop_normalize_u8: bad sizes");
  TB_REQUIRE(top >= 0 && left >= 0 && crop_h > 0 && crop_w > 0 && top + crop_h <= out_h && left + crop_w <= out_w,
             TB_E_SHAPE, "tb_resize_crop_normalize_u8: crop window [%d:%d, %d:%d] outside the %dx%d resized image",
             top, top + crop_h, left, left + crop_w, out_h, out_w);
  TB_REQUIRE(row0 >= 0 && nrows > 0 && row0 + nrows <= src_h, TB_E_SHAPE,
             "tb_resize_crop_normalize_u8: source rows [%d, %d) outside the image (%d rows)", row0, row0 + nrows,
             src_h);
  TB_REQUIRE(std != 0.f, TB_E_ARG, "tb_resize_crop_normalize_u8: std == 0");
  resample_h_u8_kernel<<<image_grid_for((long long)nrows * crop_w * channels), 256, 0, st>>>(
      (const unsigned char*)src_u8, (long long)src_w * channels, bounds_x, kk_x, ksize_x, row0, nrows, left, crop_w,
      channels, (unsigned char*)mid_u8);
  if ((rc = check_launch("resample_h_u8_kernel"))) return rc;
  resample_v_norm_kernel<<<image_grid_for((long long)crop_h * crop_w * channels), 256, 0, st>>>(
      (const unsigned char*)mid_u8, bounds_y, kk_y, ksize_y, row0, top, crop_h, crop_w, channels, scale, mean, std,
      out_f32_chw, (unsigned char*)out_u8_hwc);
  return check_launch("resample_v_norm_kernel");
}
