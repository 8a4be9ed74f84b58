// Exact image primitives of the deferred augmentation: the operations
// /root/reference/textboost/augment/paired_augmentation.py runs with PIL / torchvision on the host for every training
// item — edge padding, centre / box crops, mirror, the framed n x n collage, Pillow's AFFINE transform with bicubic or
// nearest sampling, the fixed-point grayscale — as byte-exact kernels over uint8 [H, W, C] images resident in HBM.
// Arithmetic restated from Pillow (src/libImaging/Geometry.c, Convert.c) and torchvision (center_crop, pad), pinned on
// the CPU by oracle/pil_affine_ref.py + tests/test_resample_cpu.py / test_image_plan_cpu.py.  HBM-bound byte work
// (a 1024^2 RGB image is 3 MB): coalesced interleaved-channel accesses, grids capped at a few CTAs per SM; the
// bicubic transform is ~60 double-precision operations per output byte, evaluated with explicitly rounded
// multiplies / adds so that nothing contracts into an FMA (Pillow's build does not contract).
#include "host_util.h"

namespace tb {

static inline unsigned aug_grid_for(long long n) {
  long long b = (n + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

// out[y, x, c] = src[sy, sx, c] with (sx, sy) = (x, y) folded by the tile size (collage), mirrored (flip), shifted by
// (ox, oy); outside the source: the nearest edge pixel (clamp = 1: edge padding) or zero.  frame = 1 blanks the
// outermost pixel ring of every tile.
__global__ void img_gather_u8_kernel(const unsigned char* __restrict__ src, int src_h, int src_w, int channels,
                                     unsigned char* __restrict__ out, int out_h, int out_w, int ox, int oy, int clamp,
                                     int flip_x, int tile_w, int tile_h, int frame) {
  const long long total = (long long)out_h * out_w * channels;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % channels);
    const long long t = i / channels;
    int x = (int)(t % out_w), y = (int)(t / out_w);
    bool zero = false;
    if (tile_w > 0) {
      x %= tile_w;
      y %= tile_h;
      zero = frame && (x == 0 || y == 0 || x == tile_w - 1 || y == tile_h - 1);
    }
    if (flip_x) x = out_w - 1 - x;
    int sx = x + ox, sy = y + oy;
    if (clamp) {
      sx = sx < 0 ? 0 : (sx >= src_w ? src_w - 1 : sx);
      sy = sy < 0 ? 0 : (sy >= src_h ? src_h - 1 : sy);
    } else if (sx < 0 || sx >= src_w || sy < 0 || sy >= src_h) {
      zero = true;
    }
    out[i] = zero ? (unsigned char)0 : src[((long long)sy * src_w + sx) * channels + c];
  }
}

__device__ __forceinline__ double cubic_m1(double v1, double v2, double v3, double v4, double d) {
  // Pillow BICUBIC: p1 + d (p2 + d (p3 + d p4)); the p's are exact (small integers), the Horner steps are rounded
  const double p1 = v2;
  const double p2 = -v1 + v3;
  const double p3 = 2 * (v1 - v2) + v3 - v4;
  const double p4 = -v1 + v2 - v3 + v4;
  return __dadd_rn(p1, __dmul_rn(d, __dadd_rn(p2, __dmul_rn(d, __dadd_rn(p3, __dmul_rn(d, p4))))));
}

__device__ __forceinline__ int floor_like_pillow(double v) { return v < 0.0 ? (int)floor(v) : (int)v; }

// Pillow ImagingGenericTransform(affine_transform, bicubic_filter32RGB | nearest_filter), same output size as input
__global__ void img_affine_u8_kernel(const unsigned char* __restrict__ src, int H, int W, int channels,
                                     unsigned char* __restrict__ out, double a0, double a1, double a2, double a3,
                                     double a4, double a5, int bicubic) {
  const long long total = (long long)H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)(i / W);
    const double xo = x + 0.5, yo = y + 0.5;
    double xin = __dadd_rn(__dadd_rn(__dmul_rn(a0, xo), __dmul_rn(a1, yo)), a2);
    double yin = __dadd_rn(__dadd_rn(__dmul_rn(a3, xo), __dmul_rn(a4, yo)), a5);
    unsigned char* o = out + i * channels;
    if (xin < 0.0 || xin >= W || yin < 0.0 || yin >= H) {
      for (int c = 0; c < channels; ++c) o[c] = 0;
      continue;
    }
    if (!bicubic) {
      const unsigned char* p = src + ((long long)(int)yin * W + (int)xin) * channels;
      for (int c = 0; c < channels; ++c) o[c] = p[c];
      continue;
    }
    xin -= 0.5;
    yin -= 0.5;
    int xi = floor_like_pillow(xin), yi = floor_like_pillow(yin);
    const double dx = xin - xi, dy = yin - yi;
    --xi;
    --yi;
    int col[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int cx = xi + k;
      col[k] = (cx < 0 ? 0 : (cx >= W ? W - 1 : cx)) * channels;
    }
    for (int c = 0; c < channels; ++c) {
      double v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int yy = yi + k;
        if (k > 0 && (yy < 0 || yy >= H)) {
          v[k] = v[k - 1];  // rows outside the image repeat the value above
          continue;
        }
        const int cy = yy < 0 ? 0 : (yy >= H ? H - 1 : yy);  // only the first row is clamped
        const unsigned char* row = src + (long long)cy * W * channels + c;
        v[k] = cubic_m1((double)row[col[0]], (double)row[col[1]], (double)row[col[2]], (double)row[col[3]], dx);
      }
      const double r = cubic_m1(v[0], v[1], v[2], v[3], dy);
      o[c] = r <= 0.0 ? (unsigned char)0 : (r >= 255.0 ? (unsigned char)255 : (unsigned char)(int)r);
    }
  }
}

// PIL "RGB" -> "L" -> "RGB": L = (R * 19595 + G * 38470 + B * 7471 + 0x8000) >> 16
__global__ void img_grayscale_u8_kernel(const unsigned char* __restrict__ src, unsigned char* __restrict__ out,
                                        long long npix) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix;
       i += (long long)gridDim.x * blockDim.x) {
    const unsigned char* p = src + i * 3;
    const unsigned char l = (unsigned char)(((int)p[0] * 19595 + (int)p[1] * 38470 + (int)p[2] * 7471 + 0x8000) >> 16);
    unsigned char* o = out + i * 3;
    o[0] = l;
    o[1] = l;
    o[2] = l;
  }
}

}  // namespace tb

using namespace tb;

#define TB_ENTER()            \
  int rc = tb_check_device(); \
  if (rc) return rc;          \
  cudaStream_t st = (cudaStream_t)stream

extern "C" int tb_img_gather_u8(const void* src, int src_h, int src_w, int channels, void* out, int out_h, int out_w,
                                int offset_x, int offset_y, int clamp_to_edge, int flip_x, int tile_w, int tile_h,
                                int frame, void* stream) {
  TB_ENTER();
  TB_REQUIRE(src && out && src_h > 0 && src_w > 0 && out_h > 0 && out_w > 0 && channels >= 1 && channels <= 4,
             TB_E_ARG, "tb_img_gather_u8: bad args");
  TB_REQUIRE((tile_w > 0) == (tile_h > 0), TB_E_ARG, "tb_img_gather_u8: tile_w and tile_h go together");
  img_gather_u8_kernel<<<aug_grid_for((long long)out_h * out_w * channels), 256, 0, st>>>(
      (const unsigned char*)src, src_h, src_w, channels, (unsigned char*)out, out_h, out_w, offset_x, offset_y,
      clamp_to_edge, flip_x, tile_w, tile_h, frame);
  return check_launch("img_gather_u8_kernel");
}

extern "C" int tb_img_affine_u8(const void* src, int H, int W, int channels, void* out, const double* matrix6,
                                int bicubic, void* stream) {
  TB_ENTER();
  TB_REQUIRE(src && out && matrix6 && H > 0 && W > 0 && channels >= 1 && channels <= 4, TB_E_ARG,
             "tb_img_affine_u8: bad args");
  TB_REQUIRE(src != out, TB_E_ARG, "tb_img_affine_u8: in-place transform");
  img_affine_u8_kernel<<<aug_grid_for((long long)H * W), 256, 0, st>>>(
      (const unsigned char*)src, H, W, channels, (unsigned char*)out, matrix6[0], matrix6[1], matrix6[2], matrix6[3],
      matrix6[4], matrix6[5], bicubic);
  return check_launch("img_affine_u8_kernel");
}

extern "C" int tb_img_grayscale_u8(const void* src, void* out, int64_t npix, void* stream) {
  TB_ENTER();
  TB_REQUIRE(src && out && npix > 0, TB_E_ARG, "tb_img_grayscale_u8: bad args");
  img_grayscale_u8_kernel<<<aug_grid_for(npix), 256, 0, st>>>((const unsigned char*)src, (unsigned char*)out,
                                                               (long long)npix);
  return check_launch("img_grayscale_u8_kernel");
}
