// Elementwise kernels of the validation / inference sampler (/root/reference/train_textboost.py:453-531,
// /root/reference/inference.py:84-105 through diffusers StableDiffusionPipeline + DPMSolverMultistepScheduler):
// one fused launch per denoising step for classifier-free guidance + the DPM-Solver++(2M) update + the fp16 doubled
// batch the next UNet call reads, the post_quant_conv in front of the VAE decoder, and the uint8 image write-out.
// The heavy work of a sampling step is the UNet forward (tb_conv3x3_f16 / tb_gemm_f16 / tb_attn_fwd_f16 ...).
#include "host_util.h"
#include "sm100.cuh"

namespace tb {

static inline unsigned sampler_grid_for(long long n, int threads) {
  long long b = (n + threads - 1) / threads;
  const long long cap = (long long)num_sms() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

// eps [2n] = UNet output for the [uncond | cond] doubled batch.  Per element:
//   e   = e_u + g (e_c - e_u)                                   classifier-free guidance
//   m0  = (x - s_i e) / a_i        (epsilon)   |   a_i x - s_i e   (v-prediction)       data prediction at sigma_i
//   x'  = c_x x + c_d0 m0 + c_d1 (m0 - m1)                       DPM-Solver++ first (c_d1 = 0) / second order (midpoint)
// and writes x' (fp32), m0 (fp32, next step's m1) and x' as fp16 into both halves of the next UNet input.
__global__ void dpm_cfg_step_kernel(float* __restrict__ x, const tb::half_t* __restrict__ eps,
                                    const float* __restrict__ m_prev, float* __restrict__ m_out,
                                    tb::half_t* __restrict__ unet_in, long long n, float g, float a_i, float s_i,
                                    int v_pred, float c_x, float c_d0, float c_d1) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float eu = tb::h2f(eps[i]), ec = tb::h2f(eps[n + i]);
    const float e = eu + g * (ec - eu);
    const float xi = x[i];
    const float m0 = v_pred ? a_i * xi - s_i * e : (xi - s_i * e) / a_i;
    float xn = c_x * xi + c_d0 * m0;
    if (c_d1 != 0.f) xn += c_d1 * (m0 - m_prev[i]);
    x[i] = xn;
    m_out[i] = m0;
    if (unet_in) {
      const tb::half_t h = tb::f2h(xn);
      unet_in[i] = h;
      unet_in[n + i] = h;
    }
  }
}

// z[b, o, p] = bias[o] + sum_i w[o, i] * latents[b, i, p] * inv_scale   (post_quant_conv, 1x1, L <= 8), fp16 NCHW out
__global__ void vae_decode_in_kernel(const float* __restrict__ latents, const float* __restrict__ w,
                                     const float* __restrict__ bias, tb::half_t* __restrict__ z, int L, int HW,
                                     long long total, float inv_scale) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    const long long t = i / HW;
    const int o = (int)(t % L);
    const long long b = t / L;
    float acc = bias[o];
    for (int c = 0; c < L; ++c) acc += w[o * L + c] * (latents[(b * L + c) * HW + p] * inv_scale);
    z[i] = tb::f2h(acc);
  }
}

// VaeImageProcessor.postprocess: (x / 2 + 0.5).clamp(0, 1) * 255, round half to even -> uint8 [npix, channels]
__global__ void image_u8_kernel(const tb::half_t* __restrict__ x, long long ld, unsigned char* __restrict__ out,
                                long long npix, int channels) {
  const long long total = npix * channels;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / channels;
    const int c = (int)(i - p * channels);
    const float v = fminf(fmaxf(tb::h2f(x[p * ld + c]) * 0.5f + 0.5f, 0.f), 1.f);
    out[i] = (unsigned char)rintf(v * 255.f);
  }
}

}  // namespace tb

using namespace tb;

#define TB_ENTER()            \
  int rc = tb_check_device(); \
  if (rc) return rc;          \
  cudaStream_t st = (cudaStream_t)stream

extern "C" int tb_dpm_cfg_step(float* x, const void* eps_f16, const float* m_prev, float* m_out, void* unet_in_f16,
                               int64_t n, float guidance_scale, float alpha_i, float sigma_i, int v_prediction,
                               float c_x, float c_d0, float c_d1, void* stream) {
  TB_ENTER();
  TB_REQUIRE(x && eps_f16 && m_out && n > 0, TB_E_ARG, "tb_dpm_cfg_step: bad args");
  TB_REQUIRE(c_d1 == 0.f || m_prev, TB_E_ARG, "tb_dpm_cfg_step: second-order update needs m_prev");
  TB_REQUIRE(alpha_i > 0.f, TB_E_ARG, "tb_dpm_cfg_step: alpha_i must be positive");
  dpm_cfg_step_kernel<<<sampler_grid_for(n, 256), 256, 0, st>>>(x, (const tb::half_t*)eps_f16, m_prev, m_out,
                                                                (tb::half_t*)unet_in_f16, (long long)n,
                                                                guidance_scale, alpha_i, sigma_i, v_prediction, c_x,
                                                                c_d0, c_d1);
  return check_launch("dpm_cfg_step_kernel");
}

extern "C" int tb_vae_decode_in(const float* latents, const float* w, const float* bias, void* z_f16, int B,
                                int latent_channels, int HW, float inv_scaling_factor, void* stream) {
  TB_ENTER();
  TB_REQUIRE(latents && w && bias && z_f16 && B > 0 && HW > 0 && latent_channels > 0 && latent_channels <= 8,
             TB_E_ARG, "tb_vae_decode_in: bad args");
  const long long total = (long long)B * latent_channels * HW;
  vae_decode_in_kernel<<<sampler_grid_for(total, 256), 256, 0, st>>>(latents, w, bias, (tb::half_t*)z_f16,
                                                                     latent_channels, HW, total,
                                                                     inv_scaling_factor);
  return check_launch("vae_decode_in_kernel");
}

extern "C" int tb_image_u8(const void* x_f16, int64_t ld, void* out_u8, int64_t npix, int channels, void* stream) {
  TB_ENTER();
  TB_REQUIRE(x_f16 && out_u8 && npix > 0 && channels > 0 && ld >= channels, TB_E_ARG, "tb_image_u8: bad args");
  image_u8_kernel<<<sampler_grid_for(npix * channels, 256), 256, 0, st>>>((const tb::half_t*)x_f16, (long long)ld,
                                                                         (unsigned char*)out_u8, (long long)npix,
                                                                         channels);
  return check_launch("image_u8_kernel");
}
