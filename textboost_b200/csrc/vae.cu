// HBM-bound helpers of the AutoencoderKL encoder (vae.encode at /root/reference/train_textboost.py:1036-1037):
// the right/bottom-padded stride-2 window gather of diffusers Downsample2D(padding=0), the row softmax of the
// single-head 512-channel mid-block attention (QK^T and PV run as tb_gemm_f16 calls), and the diagonal-Gaussian
// sample + scaling_factor that turns the moments into training latents.  The convolutions, GroupNorms and linears
// of the encoder reuse tb_conv3x3_f16 / tb_gemm_f16 / tb_groupnorm_fwd_f16 / tb_conv_in_f16.
#include "host_util.h"
#include "sm100.cuh"

namespace tb {

static inline unsigned vae_grid_for(long long n, int threads) {
  long long b = (n + threads - 1) / threads;
  const long long cap = (long long)num_sms() * 32;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

// col[(b,oy,ox), (ky*3+kx)*C + c] = x[b, 2*oy+ky-pad_lo, 2*ox+kx-pad_lo, c]  (zero outside the image).
// pad_lo = 1: UNet Downsample2D (symmetric pad 1); pad_lo = 0: VAE Downsample2D = F.pad(x, (0,1,0,1)) + conv pad 0.
__global__ void im2col3x3s2_pad_kernel(const tb::half_t* __restrict__ x, tb::half_t* __restrict__ col, int B, int H, int W,
                                       int C, int pad_lo) {
  const int Ho = H / 2, Wo = W / 2, cv = C / 8;
  const long long total = (long long)B * Ho * Wo * 9 * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    long long t = i / cv;
    const int tap = (int)(t % 9);
    t /= 9;
    const int ox = (int)(t % Wo);
    t /= Wo;
    const int oy = (int)(t % Ho);
    const int b = (int)(t / Ho);
    const int iy = 2 * oy + tap / 3 - pad_lo, ix = 2 * ox + tap % 3 - pad_lo;
    uint4 q = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W)
      q = *reinterpret_cast<const uint4*>(x + (((long long)b * H + iy) * W + ix) * C + v * 8);
    *reinterpret_cast<uint4*>(col + i * 8) = q;
  }
}

// In-place softmax over each row of an fp16 matrix (fp32 arithmetic).  One CTA of 256 threads per row, the row held
// in registers (cols <= 8192): one read and one write of the score matrix.
constexpr int SM_MAXV = 4;
__global__ void __launch_bounds__(256) softmax_rows_kernel(tb::half_t* __restrict__ x, long long ld, int cols) {
  __shared__ float red[8];
  __shared__ float bcast;
  tb::half_t* row = x + (long long)blockIdx.x * ld;
  const int nv = cols / 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float f[SM_MAXV][8];
  float m = -INFINITY;
#pragma unroll
  for (int j = 0; j < SM_MAXV; ++j) {
    const int v = threadIdx.x + j * 256;
    if (v < nv) {
      const uint4 q = *reinterpret_cast<const uint4*>(row + v * 8);
      const tb::half2_t* h = reinterpret_cast<const tb::half2_t*>(&q);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 t = tb::h22f2(h[i]);
        f[j][2 * i] = t.x;
        f[j][2 * i + 1] = t.y;
        m = fmaxf(m, fmaxf(t.x, t.y));
      }
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[warp] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = red[0];
    for (int w = 1; w < 8; ++w) t = fmaxf(t, red[w]);
    bcast = t;
  }
  __syncthreads();
  m = bcast;
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < SM_MAXV; ++j) {
    if (threadIdx.x + j * 256 < nv) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        f[j][i] = __expf(f[j][i] - m);
        s += f[j][i];
      }
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __syncthreads();  // everyone has read bcast
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    bcast = 1.f / t;
  }
  __syncthreads();
  const float inv = bcast;
#pragma unroll
  for (int j = 0; j < SM_MAXV; ++j) {
    const int v = threadIdx.x + j * 256;
    if (v < nv) {
      uint4 o;
      o.x = pack_half2(f[j][0] * inv, f[j][1] * inv);
      o.y = pack_half2(f[j][2] * inv, f[j][3] * inv);
      o.z = pack_half2(f[j][4] * inv, f[j][5] * inv);
      o.w = pack_half2(f[j][6] * inv, f[j][7] * inv);
      *reinterpret_cast<uint4*>(row + v * 8) = o;
    }
  }
}

// DiagonalGaussianDistribution.sample() * scaling_factor.  moments: fp16 channels-last rows [B*HW, ld], columns
// [0,L) = mean, [L,2L) = logvar (clamped to [-30, 20]); eps / latents: fp32 NCHW [B, L, HW].
__global__ void vae_sample_kernel(const tb::half_t* __restrict__ moments, long long ld, const float* __restrict__ eps,
                                  float* __restrict__ latents, float* __restrict__ mean_out,
                                  float* __restrict__ std_out, int HW, int L, long long total, float scale) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int pix = (int)(i % HW);
    const long long t = i / HW;
    const int c = (int)(t % L);
    const long long b = t / L;
    const tb::half_t* row = moments + (b * HW + pix) * ld;
    const float mean = tb::h2f(row[c]);
    const float logvar = fminf(fmaxf(tb::h2f(row[L + c]), -30.f), 20.f);
    const float sd = expf(0.5f * logvar);
    if (latents) latents[i] = (mean + sd * eps[i]) * scale;
    if (mean_out) mean_out[i] = mean;
    if (std_out) std_out[i] = sd;
  }
}

}  // namespace tb

using namespace tb;

#define TB_ENTER()            \
  int rc = tb_check_device(); \
  if (rc) return rc;          \
  cudaStream_t st = (cudaStream_t)stream

extern "C" int tb_im2col3x3s2_pad_f16(const void* x, void* col, int B, int H, int W, int C, int pad_lo,
                                      void* stream) {
  TB_ENTER();
  TB_REQUIRE(x && col && C % 8 == 0 && H % 2 == 0 && W % 2 == 0 && (pad_lo == 0 || pad_lo == 1), TB_E_ARG,
             "tb_im2col3x3s2_pad_f16: bad args");
  im2col3x3s2_pad_kernel<<<vae_grid_for((long long)B * (H / 2) * (W / 2) * 9 * (C / 8), 256), 256, 0, st>>>(
      (const tb::half_t*)x, (tb::half_t*)col, B, H, W, C, pad_lo);
  return check_launch("im2col3x3s2_pad_kernel");
}

extern "C" int tb_softmax_rows_f16(void* x, int64_t ld, int64_t rows, int cols, void* stream) {
  TB_ENTER();
  TB_REQUIRE(x && rows > 0 && rows <= 0x7fffffff, TB_E_ARG, "tb_softmax_rows_f16: bad args");
  TB_REQUIRE(cols % 8 == 0 && cols > 0 && cols <= SM_MAXV * 256 * 8 && ld % 8 == 0, TB_E_SHAPE,
             "tb_softmax_rows_f16: cols=%d unsupported (multiple of 8, <= %d)", cols, SM_MAXV * 256 * 8);
  softmax_rows_kernel<<<(unsigned)rows, 256, 0, st>>>((tb::half_t*)x, (long long)ld, cols);
  return check_launch("softmax_rows_kernel");
}

extern "C" int tb_vae_sample(const void* moments_f16, int64_t ld, const float* eps, float* latents, float* mean,
                             float* std, int B, int HW, int latent_channels, float scaling_factor, void* stream) {
  TB_ENTER();
  TB_REQUIRE(moments_f16 && B > 0 && HW > 0 && latent_channels > 0 && ld >= 2 * latent_channels, TB_E_ARG,
             "tb_vae_sample: bad args");
  TB_REQUIRE((latents == nullptr) || eps, TB_E_ARG, "tb_vae_sample: latents requested without eps");
  const long long total = (long long)B * latent_channels * HW;
  vae_sample_kernel<<<vae_grid_for(total, 256), 256, 0, st>>>((const tb::half_t*)moments_f16, (long long)ld, eps,
                                                              latents, mean, std, HW, latent_channels, total,
                                                              scaling_factor);
  return check_launch("vae_sample_kernel");
}
