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
  const void* residual; /* fp16 [M, ldr] or NULL */
  int64_t ldr;
  float alpha;
  int32_t act;
  int32_t out_kind;
} tb_epilogue;

int tb_version(void);
const char* tb_last_error(void);
/* 0 if the current device is sm_100 and the library can run, else TB_E_ARCH */
int tb_check_device(void);

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
 * softmax(Q K^T * scale) V per (batch, head); no mask, no dropout: diffusers AttnProcessor2_0 ->
 * F.scaled_dot_product_attention inside unet(...) at train_textboost.py:1063 (forward) and its
 * autograd backward at :1108.  q/k/v/o are token-major fp16 [B, N, ld*] with head h occupying columns
 * [h*d, (h+1)*d) of each row, so they can point into the fused QKV projection output.
 * lse [B, heads, Nq] fp32 (log2 domain) is written by the forward and read by the backward. */
int tb_attn_fwd_f16(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                    void* o, int64_t ldo, float* lse, int B, int heads, int Nq, int Nk, int d,
                    float scale, void* stream);
/* delta: workspace [B, heads, Nq] fp32.  dQacc: fp32 [B*Nq, lddq] accumulator, zeroed here and filled
 * with red.add (NULL = dQ not needed: the first cross-attention of the UNet).
 * dK/dV fp16 [B, Nk, ldd*] head-major columns like k/v. */
int tb_attn_bwd_f16(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                    const void* o, int64_t ldo, const void* dO, int64_t lddo, const float* lse,
                    float* delta, float* dQacc, int64_t lddq, void* dK, int64_t lddk, void* dV,
                    int64_t lddv, int B, int heads, int Nq, int Nk, int d, float scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TEXTBOOST_B200_H_ */
