/* textboost_b200 — C ABI of libtextboost_b200.so
 *
 * Drop-in boundary for ONE hot path of nahyeonkaty/textboost: the TextBoost training step
 * (reference: train_textboost.py:1041-1149).  The reference is pure Python; the arithmetic it
 * triggers lives in diffusers 0.29 (UNet2DConditionModel), transformers (CLIPTextModel), peft
 * (LoRA Linear) and torch (AdamW, autograd).  Each entry point below replaces the library
 * kernel(s) that one of those calls dispatches to; the comment on each cites the reference call
 * site whose arithmetic it carries.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes, no C++/torch types.
 *  - every pointer is a DEVICE pointer owned by the caller unless it says "host".
 *  - every function enqueues on `stream` (a cudaStream_t passed as void*) and never synchronises,
 *    never allocates device memory: it is CUDA-graph capturable.
 *  - returns 0 on success, a negative TB_E_* code otherwise; tb_last_error() gives the text.
 *  - activations are fp16, channels-last (NHWC == [B, H*W, C] tokens); accumulation is fp32.
 *  - there is no CPU fallback: on a device that is not sm_100 every compute call returns TB_E_ARCH.
 */
#ifndef TEXTBOOST_B200_H_
#define TEXTBOOST_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TB_OK 0
#define TB_E_SHAPE (-1)
#define TB_E_ALIGN (-2)
#define TB_E_ARCH (-3)
#define TB_E_CUDA (-4)
#define TB_E_ARG (-5)

#define TB_ABI_VERSION 1

/* activation applied to (alpha*acc + bias + rowvec) before the residual is added */
#define TB_ACT_NONE 0
#define TB_ACT_SILU 1
#define TB_ACT_QUICK_GELU 2 /* x*sigmoid(1.702x): CLIP-L hidden_act (transformers modeling_clip) */
#define TB_ACT_GELU 3       /* erf gelu: OpenCLIP-H hidden_act */

/* output kinds */
#define TB_OUT_F16 0     /* C fp16 = result */
#define TB_OUT_F32 1     /* C fp32 = result */
#define TB_OUT_F32_ACC 2 /* C fp32 += result (d ehs accumulation over the 16 cross-attention blocks) */

/* Fused GEMM epilogue:  out = act(alpha*acc + bias[n] + rowvec[m / rows_per_group, n]) + residual[m, n] */
typedef struct tb_epilogue {
  const void* bias;     /* fp16 [N] or NULL */
  const void* rowvec;   /* fp16 [ceil(M/rows_per_group), N] or NULL (ResnetBlock2D time-embedding add) */
  int32_t rows_per_group;
  const void* residual; /* [M, ldr] fp16 (fp32 when residual_f32 != 0) or NULL */
  int64_t ldr;
  int32_t residual_f32; /* CLIP residual stream is fp32 (accelerate autocast keeps LN / adds in fp32) */
  float alpha;
  int32_t act;
  int32_t out_kind;
  int64_t ld_rowvec;    /* row stride of rowvec in elements; 0 = N.  > N: rowvec is a column slice of a wider matrix
                           (all 22 time_emb_proj of a UNet forward are ONE GEMM, each ResnetBlock2D reads its slice) */
} tb_epilogue;

int tb_version(void);
const char* tb_last_error(void);
/* 0 if the current device is sm_100 and the library can run, else TB_E_ARCH */
int tb_check_device(void);
/* Storage / tensor-core operand type of THIS build of the library: TB_STORAGE_F16 (libtextboost_b200.so, the
 * reference's --mixed_precision fp16, weight_dtype = torch.float16, train_textboost.py:928-931) or TB_STORAGE_BF16
 * (libtextboost_b200_bf16.so, --mixed_precision bf16, :932-933).  Both builds export the same symbols: wherever a
 * signature or comment below says "fp16" the bf16 build reads and writes bfloat16 instead; fp32 arguments are fp32
 * in both.  A process uses ONE of the two (textboost_b200.precision). */
#define TB_STORAGE_F16 0
#define TB_STORAGE_BF16 1
int tb_storage_dtype(void);

/* Optional split-K scratch for under-filled GEMM / conv problems (few output tiles, long K): `ptr` is a
 * caller-owned device buffer (256-byte aligned, >= 128 KiB; 32 MiB covers every SD shape) that tb_gemm_f16 /
 * tb_conv3x3_f16 calls enqueued on `stream` may use between their own start and end.  One region per stream:
 * calls on different streams can run concurrently.  ptr == NULL unregisters.  Without a workspace the kernels
 * never split (same results up to fp32 summation order).  The library still never allocates. */
int tb_set_workspace(void* stream, void* ptr, size_t bytes);

/* ---- dense contraction on tcgen05/TMEM -------------------------------------------------------
 * C[M,N] = epilogue(A[M,K] * B[N,K]^T).  A, B fp16 row-major (K contiguous; lda/ldb in elements,
 * multiples of 8).  Replaces every nn.Linear / 1x1 conv forward and input-gradient (dgrad uses the
 * pre-transposed weight) in UNet2DConditionModel and CLIPTextModel
 * (train_textboost.py:1063-1067 forward, :1108 backward).  K must be a multiple of 8. */
int tb_gemm_f16(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, int M,
                int N, int K, const tb_epilogue* ep, void* stream);

/* 3x3 convolution, stride 1, pad 1, as an implicit GEMM (TMA gathers the shifted NHWC window with
 * zero fill at the borders; no im2col buffer).  x [B,H,W,Cin] fp16, w [Cout, 9*Cin] fp16 with
 * k = (ky*3+kx)*Cin + ci, y [B,H,W,Cout].  Cin % 64 == 0.  The input-gradient of a conv3x3 is the
 * same call with the tap-flipped, in/out-transposed weight.  Replaces the cuDNN conv3x3 kernels
 * under diffusers ResnetBlock2D / Downsample2D / Upsample2D (train_textboost.py:1063, :1108). */
int tb_conv3x3_f16(const void* x, const void* w, void* y, int B, int H, int W, int Cin, int Cout,
                   const tb_epilogue* ep, void* stream);

/* ---- UNet attention (flash, tcgen05/TMEM) ------------------------------------------------------
 * softmax(Q K^T * scale) V per (batch, head); no dropout; causal=0: no mask (diffusers AttnProcessor2_0
 * -> F.scaled_dot_product_attention inside unet(...) at train_textboost.py:1063, backward at :1108);
 * causal=1 (Nq == Nk): key j visible to query i iff j <= i (transformers CLIPAttention under
 * CLIPTextTransformer's causal mask, called from textboost/text_encoder.py:62-69).  q/k/v/o are token-major fp16 [B, N, ld*] with head h occupying columns
 * [h*d, (h+1)*d) of each row, so they can point into the fused QKV projection output.
 * lse [B, heads, Nq] fp32 (log2 domain) is written by the forward and read by the backward. */
int tb_attn_fwd_f16(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                    void* o, int64_t ldo, float* lse, int B, int heads, int Nq, int Nk, int d,
                    float scale, int causal, void* stream);
/* delta: workspace [B, heads, Nq] fp32.  dQacc: fp32 [B*Nq, lddq] accumulator, zeroed here and filled
 * with red.add (NULL = dQ not needed: the first cross-attention of the UNet).
 * dQ16 (instead of dQacc, Nk <= 128 only: the 77-token cross attention and the CLIP self attention): with a single
 * KV tile each dQ row is complete inside one CTA and is stored once as fp16 [B*Nq, lddq16] -- no accumulator, no
 * memset, no cast afterwards.
 * dK/dV fp16 [B, Nk, ldd*] head-major columns like k/v. */
int tb_attn_bwd_f16(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                    const void* o, int64_t ldo, const void* dO, int64_t lddo, const float* lse,
                    float* delta, float* dQacc, int64_t lddq, void* dQ16, int64_t lddq16, void* dK,
                    int64_t lddk, void* dV, int64_t lddv, int B, int heads, int Nq, int Nk, int d, float scale,
                    int causal, void* stream);

/* debug only: buf = device int64[16*8*10] receiving clock64 timestamps of CTA (0,0,0) of the following
 * tb_attn_fwd_f16 launches ([iteration][event][warp]); NULL switches it off.  Not used by the product path. */
int tb_attn_debug_trace(void* buf);

/* ---- normalisation (HBM-bound, channels-last fp16, fp32 statistics) ---------------------------
 * GroupNorm(+SiLU): diffusers ResnetBlock2D.norm1/norm2 + nonlinearity, Transformer2DModel.norm,
 * conv_norm_out + conv_act (inside unet(...), train_textboost.py:1063 / backward :1108).
 * x,y,dy,dx [B, HW, C] fp16; gamma/beta fp16 [C]; stats [B,G,2] fp32 = (sum x, sum x^2) written by the
 * forward and consumed by the backward; dstats [B,G,2] fp32 workspace. */
/* `silu` is a flag word: bit 0 = fuse SiLU; TB_GN_STATS_ZEROED = the stats / dstats buffer is already zero (the caller
 * cleared one buffer for all 61 GroupNorms of a UNet pass and hands out slices), so no memset node is enqueued. */
#define TB_GN_STATS_ZEROED 2
int tb_groupnorm_fwd_f16(const void* x, const void* gamma, const void* beta, void* y, float* stats, int B,
                         int HW, int C, int G, float eps, int silu, void* stream);
/* dx = GN_backward(dy) + add  (add fp16 [B,HW,C] or NULL: the residual branch's gradient). */
int tb_groupnorm_bwd_f16(const void* dy, const void* x, const void* gamma, const void* beta,
                         const float* stats, float* dstats, const void* add, void* dx, int B, int HW,
                         int C, int G, float eps, int silu, void* stream);
/* LayerNorm over the last dim of [M, C].  Type sets: UNet (x, gamma/beta, y all fp16) and CLIP (x and
 * gamma/beta fp32 — the residual stream stays fp32 under accelerate's autocast — y fp16 for the next
 * GEMM or fp32 for the final LayerNorm).  stats [M,2] = (mean, rstd).  ld* in elements. */
int tb_layernorm_fwd(const void* x, int x_f32, int64_t ldx, const void* gamma, const void* beta, int w_f32,
                     void* y, int y_f32, int64_t ldy, float* stats, int M, int C, float eps, void* stream);
/* dx = LN_backward(dy) + add (add may be NULL or alias dx); dx/add/x/gamma share a type (fp16 | fp32). */
int tb_layernorm_bwd(const void* dy, int dy_f32, int64_t lddy, const void* x, int x_f32, int64_t ldx,
                     const void* gamma, const float* stats, const void* add, void* dx, int M, int C,
                     void* stream);
/* Text-encoder LayerNorm with the LoRA glue folded in (the CLIP layers run at M = batch x 77 rows, where every launch is
 * latency): replaces transformers CLIPEncoderLayer.layer_norm1 followed by peft's lora_A projection
 * (train_textboost.py:702-709 adapter, textboost/text_encoder.py:62-69 encoder call).
 * forward: y_ext[m, :C] = fp16(LN(x[m])), y_ext[m, C + j] = fp16(sum_c y[m,c] * lora_A[j,c]) for j < R and 0 for
 *          R <= j < RPAD -- the [LN(x) | LN(x) A^T] operand of the fused QKV GEMM.  x / gamma / beta / lora_A fp32. */
int tb_layernorm_lora_fwd(const float* x, int64_t ldx, const float* gamma, const float* beta, void* y_ext,
                          int64_t ldy, float* stats, const float* lora_A, int R, int RPAD, int M, int C, float eps,
                          void* stream);
/* backward (fp32 residual stream): dy fp16 [M, lddy] (fp32 when dy_f32) ; with lora_A (fp32 [R, C]) the extension
 * columns dy[m, C + j] are the gradient of the LoRA down-projection and dy_eff = dy[:, :C] + dy[:, C:C+R] lora_A;
 * dx = LN_backward(dy_eff) + add (add NULL or aliasing dx); dx_f16 (NULL or fp16 [M, C]) receives fp16(dx), the A
 * operand of the next input-gradient GEMM. */
int tb_layernorm_bwd_clip(const void* dy, int dy_f32, int64_t lddy, const float* x, int64_t ldx, const float* gamma,
                          const float* stats, const float* add, float* dx, void* dx_f16, const float* lora_A, int R,
                          int M, int C, void* stream);

/* ---- elementwise / data movement (HBM-bound) -------------------------------------------------- */
/* diffusers GEGLU: out[M,F] = h[:, :F] * gelu(h[:, F:]);  backward gives dh [M,2F]. */
int tb_geglu_fwd_f16(const void* h, void* out, int64_t M, int F, void* stream);
int tb_geglu_bwd_f16(const void* dg, const void* h, void* dh, int64_t M, int F, void* stream);
/* Upsample2D: nearest x2 on NHWC, and its adjoint (2x2 sum). */
int tb_upsample2x_fwd_f16(const void* x, void* y, int B, int H, int W, int C, void* stream);
int tb_upsample2x_bwd_f16(const void* dy, void* dx, int B, int H, int W, int C, void* stream);
/* dst[r, :cols] (=|+=) src[r, :cols] with row strides: skip-connection concat / split / residual adds. */
int tb_copy2d_f16(void* dst, int64_t ldd, const void* src, int64_t lds, int64_t rows, int cols,
                  int accumulate, void* stream);
/* dst[r, :] = [a[r, :Ca] | b[r, :Cb]] for fp16 rows with free row strides: torch.cat([hidden_states, res_hidden_states],
 * dim=1) of diffusers' up blocks on channels-last rows, one launch. */
int tb_concat2_f16(void* dst, int64_t ldd, const void* a, int64_t lda, int Ca, const void* b, int64_t ldb, int Cb,
                   int64_t rows, void* stream);
int tb_cast_f32_f16(void* dst, int64_t ldd, const void* src, int64_t lds, int64_t rows, int cols,
                    float scale, void* stream);
/* Downsample2D (conv3x3 stride 2 pad 1): forward = im2col + tb_gemm_f16; backward = zero-stuff +
 * tb_conv3x3_f16 with the flipped weight. */
int tb_im2col3x3s2_f16(const void* x, void* col, int B, int H, int W, int C, void* stream);
int tb_zero_stuff2x_f16(const void* dy, void* out, int B, int Ho, int Wo, int C, void* stream);
/* diffusers get_timestep_embedding(flip_sin_to_cos=True, freq_shift=0) -> fp16 [B, dim] = [cos | sin]. */
int tb_timestep_embedding_f16(const int64_t* t, void* out, int B, int dim, void* stream);
int tb_silu_f16(const void* x, void* y, int64_t n, void* stream);
/* dst[i] += alpha * src[i], fp32 (loss += knowledge-preservation term, train_textboost.py:1106), and a stream-ordered
 * clear of a buffer (cudaMemsetAsync: a memset node in the captured step, not a kernel) for the accumulators the step
 * starts from zero (loss, d(text states), GroupNorm sums). */
int tb_axpy_f32(float* dst, const float* src, int64_t n, float alpha, void* stream);
int tb_fill_zero(void* ptr, size_t bytes, void* stream);
/* DDPMScheduler.add_noise (+ target: epsilon or get_velocity), train_textboost.py:1052, :1070-1075.
 * x0/eps fp32 [B, per_image]; acp = alphas_cumprod fp32 [T]; noisy fp16; target fp32 (may be NULL). */
int tb_add_noise(const float* x0, const float* eps, const int64_t* t, const float* acp, void* noisy_f16,
                 float* target, int B, int per_image, int v_prediction, void* stream);
/* F.mse_loss(pred.float(), target.float()).mean() fused with its gradient (train_textboost.py:1085-1090):
 * *loss_acc += weight*mse;  dpred = weight * 2 (pred-target)/n * (*loss_scale). */
int tb_mse_fwd_bwd(const void* pred_f16, const float* target, int64_t n, float weight,
                   const float* loss_scale, float* loss_acc, void* dpred_f16, void* stream);
/* conv_in (4 -> C0) reads the NCHW fp16 sample and writes NHWC; conv_out (C0 -> 4) does the reverse.
 * Weights in the diffusers layout [Cout, Cin, 3, 3]. */
int tb_conv_in_f16(const void* x_nchw, const void* w, const void* bias, void* y_nhwc, int B, int H, int W,
                   int Cin, int Cout, void* stream);
int tb_conv_out_f16(const void* h_nhwc, const void* w, const void* bias, void* y_nchw, int B, int H, int W,
                    int Cin, int Cout, void* stream);
int tb_conv_out_bwd_f16(const void* dy_nchw, const void* w, void* dh_nhwc, int B, int H, int W, int Cin,
                        int Cout, void* stream);

/* ---- AutoencoderKL encoder helpers (vae.encode(pixel_values).latent_dist.sample() * scaling_factor,
 * train_textboost.py:1036-1037; the encoder's convolutions / GroupNorms / linears are tb_conv_in_f16,
 * tb_conv3x3_f16, tb_gemm_f16 and tb_groupnorm_fwd_f16 calls). ---- */
/* tb_im2col3x3s2_f16 with a selectable low-side padding: pad_lo = 1 is the UNet's Downsample2D, pad_lo = 0 the
 * VAE's (F.pad(x, (0,1,0,1)) then conv stride 2 pad 0: zeros on the right / bottom edge only). */
int tb_im2col3x3s2_pad_f16(const void* x, void* col, int B, int H, int W, int C, int pad_lo, void* stream);
/* In-place row softmax of an fp16 [rows, cols] matrix with row stride ld (fp32 arithmetic): the probabilities of
 * the single-head mid-block attention between its QK^T and PV GEMMs.  cols % 8 == 0, cols <= 8192. */
int tb_softmax_rows_f16(void* x, int64_t ld, int64_t rows, int cols, void* stream);
/* DiagonalGaussianDistribution: moments fp16 channels-last [B*HW, ld] (columns [0,L) mean, [L,2L) logvar, clamped
 * to [-30,20]); eps fp32 NCHW [B,L,HW].  latents = (mean + exp(logvar/2) * eps) * scaling_factor (fp32 NCHW);
 * mean / std are optional fp32 NCHW outputs; latents may be NULL when only the moments are wanted. */
int tb_vae_sample(const void* moments_f16, int64_t ld, const float* eps, float* latents, float* mean, float* std,
                  int B, int HW, int latent_channels, float scaling_factor, void* stream);

/* ---- validation / inference sampler (diffusers StableDiffusionPipeline + DPMSolverMultistepScheduler, reached from
 * train_textboost.py:453-531 log_validation and inference.py:84-105).  The UNet forward of every denoising step is the
 * tb_gemm_f16 / tb_conv3x3_f16 / tb_attn_fwd_f16 ... calls of the training path; these are the steps in between. ---- */
/* One fused launch per denoising step: classifier-free guidance e = e_u + g (e_c - e_u) on the UNet output of the
 * [uncond | cond] doubled batch (eps fp16 [2n]); data prediction m0 at (alpha_i, sigma_i) for epsilon or
 * v-prediction; DPM-Solver++ update x <- c_x x + c_d0 m0 + c_d1 (m0 - m_prev)  (c_d1 = 0: first order; midpoint
 * second order otherwise; the coefficients are host scalars of the sigma schedule).  x fp32 [n] in place, m_out fp32
 * [n], and x as fp16 into both halves of unet_in [2n] (may be NULL on the last step). */
int tb_dpm_cfg_step(float* x, const void* eps_f16, const float* m_prev, float* m_out, void* unet_in_f16, int64_t n,
                    float guidance_scale, float alpha_i, float sigma_i, int v_prediction, float c_x, float c_d0,
                    float c_d1, void* stream);
/* vae.decode input: z = post_quant_conv(latents / scaling_factor), 1x1 conv with w fp32 [L,L], bias fp32 [L];
 * latents fp32 NCHW [B,L,HW] -> z fp16 NCHW (read by tb_conv_in_f16 of the decoder). */
int tb_vae_decode_in(const float* latents, const float* w, const float* bias, void* z_f16, int B,
                     int latent_channels, int HW, float inv_scaling_factor, void* stream);
/* VaeImageProcessor.postprocess: (x/2 + 0.5).clamp(0,1) * 255 rounded half-to-even -> uint8 [npix, channels] from fp16
 * channels-last rows with stride ld. */
int tb_image_u8(const void* x_f16, int64_t ld, void* out_u8, int64_t npix, int channels, void* stream);

/* ---- image front end (TextBoostDataset's Resize(LANCZOS) -> crop -> ToImage/ToDtype/Normalize,
 * /root/reference/textboost/dataset.py:326-351, 386-388), byte-exact on a decoded uint8 image resident in HBM. ---- */
/* Pillow's antialiased 8-bit resize (src/libImaging/Resample.c: horizontal then vertical pass, 22-bit fixed-point
 * weights, uint8 intermediate) of src [src_h, src_w, channels] to out_w x out_h, restricted to the crop window
 * [top, top+crop_h) x [left, left+crop_w) of the resized image, then out = ((float)byte * scale - mean) / std as fp32
 * [channels, crop_h, crop_w] (scale = (float)(1/255.0): torchvision ToDtype; mean = std = 0.5: Normalize) and / or the
 * cropped bytes [crop_h, crop_w, channels].  bounds_* int32 [out, 2] = (first input index, tap count) and kk_* int32
 * [out, ksize] are the weight tables of each axis, computed on the host in double precision exactly as Pillow's
 * precompute_coeffs / normalize_coeffs_8bpc.  [row0, row0+nrows) = source rows the crop window's vertical taps touch;
 * mid_u8 = caller-owned scratch of nrows * crop_w * channels bytes. */
int tb_resize_crop_normalize_u8(const void* src_u8, int src_h, int src_w, int channels, const int32_t* bounds_x,
                                const int32_t* kk_x, int ksize_x, int out_w, const int32_t* bounds_y,
                                const int32_t* kk_y, int ksize_y, int out_h, int row0, int nrows, int top, int left,
                                int crop_h, int crop_w, float scale, float mean, float std, void* mid_u8,
                                float* out_f32_chw, void* out_u8_hwc, void* stream);

/* ---- exact image primitives of the deferred augmentation (PIL / torchvision calls inside
 * /root/reference/textboost/augment/paired_augmentation.py), uint8 [H, W, channels] images in HBM. ---- */
/* out[y,x] = src[(y mod tile_h) + offset_y, mirror((x mod tile_w)) + offset_x]; outside the source: nearest edge pixel
 * (clamp_to_edge: v2.functional.pad(..., "edge")) or zero (crop / center_crop padding); frame = 1 blanks the outer
 * pixel ring of every tile (square_photo_collage, :236-260); tile_w = tile_h = 0: no tiling. */
int tb_img_gather_u8(const void* src, int src_h, int src_w, int channels, void* out, int out_h, int out_w,
                     int offset_x, int offset_y, int clamp_to_edge, int flip_x, int tile_w, int tile_h, int frame,
                     void* stream);
/* Pillow Image.transform(size, AFFINE, matrix6, BICUBIC | NEAREST) at the input size (v2.functional.affine in
 * adjust_scale :27-31 and horizontal_translate :117-123): double-precision coordinates, a = -1 cubic over the clamped
 * 4x4 neighbourhood, truncating store, zero outside.  matrix6 is a HOST pointer to the six inverse-affine doubles. */
int tb_img_affine_u8(const void* src, int H, int W, int channels, void* out, const double* matrix6, int bicubic,
                     void* stream);
/* PIL.ImageOps.grayscale(img).convert("RGB") (:155): L = (19595 R + 38470 G + 7471 B + 0x8000) >> 16 on RGB triples. */
int tb_img_grayscale_u8(const void* src, void* out, int64_t npix, void* stream);

/* ---- CLIP text encoder pieces that are not GEMM / LayerNorm ------------------------------------
 * (transformers CLIPTextTransformer called from textboost/text_encoder.py:62-69; peft LoRA Linear
 * configured at train_textboost.py:702-709.) */
/* x[m,:] = tok(ids[m]) + pos[m % L]; tok(id) = id < n_base ? base[id] * (*decay) : added[id - n_base].
 * An id outside [0, n_base + n_rows) -- torch's embedding lookup raises for it -- fills its row with NaN instead of
 * dereferencing (the loss turns NaN and the GradScaler skips the step). */
int tb_clip_embed(const int64_t* ids, const float* base, const float* added, const float* decay,
                  const float* pos, float* x, int M, int L, int D, int n_base, int n_rows, void* stream);
/* grad_rows[id - n_base, :] += g[m, :] for id >= n_base only (train_textboost.py:1109-1117). */
int tb_clip_embed_grad(const int64_t* ids, const float* g, float* grad_rows, int M, int D, int n_base,
                       void* stream);
/* peft LoRA Linear (tuners/lora/layer.py, y = W x + b + (alpha/r) B A x; configured at train_textboost.py:702-709)
 * fused into a projection GEMM as a K-extension.  A fused weight stacks nblk row blocks of D outputs (q | k | v:
 * nblk 3; out_proj: nblk 1); bit b of tmask says block b carries LoRA; T = popcount(tmask), rank r <= 16,
 * R = T*r <= RPAD <= 64 extension columns.
 * y_ext[:, D:D+R] = y_ext[:, :D] @ A^T (A fp32 [R, D]); columns D+R..D+RPAD-1 are zeroed. */
int tb_lora_down(void* y_ext, int64_t ld, const float* A, int M, int D, int R, int RPAD, void* stream);
/* write scaling*B (fp32 [T][D][r]) into the extension block of Wext [nblk*D, D+RPAD] and WextT [D+RPAD, nblk*D]. */
int tb_lora_pack(const float* B, void* Wext, void* WextT, int nblk, int tmask, int D, int r, int RPAD,
                 float scaling, void* stream);
/* dB [T][D][r] += scaling * dY^T xa ;  dA [T*r][D] += dxa^T y   (accumulating, fp32); dY is [M, nblk*D]. */
int tb_lora_grad(const void* dY, const void* y_ext, const void* dA_ext, int64_t ld, float* dB, float* dA,
                 int M, int nblk, int tmask, int D, int r, float scaling, void* stream);
/* dA_ext[:, :D] += dA_ext[:, D:D+R] @ A. */
int tb_lora_dx(void* dA_ext, int64_t ld, const float* A, int M, int D, int R, void* stream);
/* ---- UNet cross-attention K/V LoRA (--unet_params_to_train crossattn_kv, train_textboost.py:712-721, 838-841: peft
 * adapters on attn2.to_k / attn2.to_v of every transformer block, third optimiser group).  The blocks' K | V
 * projections are one fused GEMM over the text states (KV output columns); the n_adapters adapters are stacked:
 * A fp32 [n_adapters*r, ctx], B fp32 [KV, r] (row j = lora_B row of fused output column j), blk int32 [KV] = adapter of
 * column j, off int32 [n_adapters+1] = column range of each adapter (multiples of 8: a 16-byte vector of K/V columns
 * never straddles two adapters; KV % 8 == 0).
 * fwd: Z (fp32 [M, n_adapters*r], kept for the backward) = ehs A^T;  kv[m, j] += scaling * Z[m, blk[j] r ..] . B[j, :]
 * bwd: dZ (fp32 scratch [M, n_adapters*r]) = scaling * dkv B per adapter;  dB += scaling * dkv^T Z;  dA += dZ^T ehs;
 *      d_ehs (fp32 [M, ctx]) += dZ A.  ehs, kv, dkv are 16-bit [M, ctx] / [M, KV] contiguous. */
int tb_unet_lora_fwd(const void* ehs, const float* A, const float* B, const int32_t* blk, float* Z, void* kv, int M,
                     int ctx, int KV, int n_adapters, int r, float scaling, void* stream);
int tb_unet_lora_bwd(const void* dkv, const void* ehs, const float* A, const float* B, const float* Z,
                     const int32_t* blk, const int32_t* off, float* dZ, float* dA, float* dB, float* d_ehs, int M,
                     int ctx, int KV, int n_adapters, int r, float scaling, void* stream);
/* causal attention over L <= 128 tokens, head_dim 64; qkv [B*L, 3D] fused; dqkv laid out like qkv. */
int tb_clip_attn_fwd(const void* qkv, void* out, int B, int L, int D, int heads, void* stream);
int tb_clip_attn_bwd(const void* qkv, const void* dO, void* dqkv, int B, int L, int D, int heads,
                     void* stream);
int tb_act_fwd_f16(const void* u, void* out, int64_t n, int kind, void* stream);
int tb_act_bwd_f16(const void* u, const void* g, void* out, int64_t n, int kind, void* stream);
/* TextBoostModel.forward override (textboost/text_encoder.py:71-86): rows whose ids[b,1]==eos_id become
 * null_emb [L,D]; with use_fixed, position 0 of every row becomes null_emb[0].  backward=1 zeroes the
 * gradient at the overwritten slots instead. */
int tb_null_override(const int64_t* ids, const float* null_emb, float* h, int B, int L, int D, int eos_id,
                     int use_fixed, int backward, void* stream);
/* knowledge-preservation loss (train_textboost.py:1096-1106): kind 0 = 1-cos, 1 = mse, mean over the M
 * rows; *loss_acc += weight*kp; dh_acc += d(weight*kp)/dh * (*loss_scale)  (dh_acc may be NULL). */
int tb_kpl_fwd_bwd(const float* h, const float* h0, int M, int D, int kind, float weight,
                   const float* loss_scale, float* loss_acc, float* dh_acc, void* stream);

/* ---- optimiser tail (train_textboost.py:1109-1149; torch.optim.AdamW; accelerate GradScaler) ----
 * Flat fp32 buffers [LoRA (n_lora) | added embedding rows (n_rows*D)].  state: fp32[16] on the device:
 * [0] loss scale [1] growth tracker [2] found_inf [3] sum g^2 [4] step [5] frozen-row decay
 * [6] clip coef [7] grad norm [8] skipped steps [9] learning-rate multiplier of the last step.
 * grads are consumed AND zeroed.  lr_schedule (TB_LR_*) restates diffusers.optimization.get_scheduler
 * (train_textboost.py:911-916) on the device: both learning rates are multiplied by the schedule's value at the
 * number of successful steps so far (state[4]), so the schedule advances inside a replayed CUDA graph and pauses on
 * GradScaler-skipped steps, as accelerate's scheduler wrapper does. */
#define TB_LR_CONSTANT 0
#define TB_LR_CONSTANT_WITH_WARMUP 1
#define TB_LR_LINEAR 2
#define TB_LR_COSINE 3
#define TB_LR_COSINE_WITH_RESTARTS 4
#define TB_LR_POLYNOMIAL 5
int tb_optim_mix_mask(float* grad_lora_b, int64_t n, int D, int r, int parity, void* stream);
int tb_adamw_fused_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n_lora,
                        int n_rows, int D, float lr_lora, float lr_emb, float beta1, float beta2, float eps,
                        float weight_decay, float max_grad_norm, float inv_world, float mean_norm,
                        int lr_schedule, float lr_warmup_steps, float lr_total_steps,
                        float* state, float* row_norm_mean, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TEXTBOOST_B200_H_ */
