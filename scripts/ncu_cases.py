"""Single launches of the step's representative kernels at BASELINE shapes (SD-1.5, B=8), for

  ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/rNN_cases \
      python scripts/ncu_cases.py [case ...]

Each case runs twice untimed, then once between cudaProfilerStart/Stop.  Numbers printed under ncu are not
bench values.  Cases: every case_* function below (gemm_*, conv_*, gn*, attn*, clip_glue, ln, geglu)
"""
import sys

import torch

sys.path.insert(0, ".")
from textboost_b200 import _cabi as C, ops  # noqa: E402

dev = "cuda"
torch.manual_seed(0)


def rnd(*shape, dtype=torch.float16, s=1.0):
    return (torch.randn(*shape, device=dev) * s).to(dtype)


def case_gemm_k320():  # attn out-proj / proj_out at 64x64: HBM-bound
    a, w, b, r = rnd(32768, 320), rnd(320, 320, s=0.05), rnd(320), rnd(32768, 320)
    return lambda: ops.gemm(a, w, bias=b, residual=r)


def case_gemm_m2048():
    a, w, b, r = rnd(2048, 1280), rnd(1280, 1280, s=0.03), rnd(1280), rnd(2048, 1280)
    return lambda: ops.gemm(a, w, bias=b, residual=r)


def case_gemm_ff():  # GEGLU in-projection at 64x64
    a, w, b = rnd(32768, 320), rnd(2560, 320, s=0.05), rnd(2560)
    return lambda: ops.gemm(a, w, bias=b)


def case_gemm_clip():  # CLIP fc2, fp32 residual stream
    a, w, b = rnd(1232, 3072), rnd(768, 3072, s=0.02), rnd(768)
    r = rnd(1232, 768, dtype=torch.float32)
    return lambda: ops.gemm(a, w, bias=b, residual=r, out_kind=C.TB_OUT_F32)


def case_conv_h8():
    x, w, b = rnd(8, 8, 8, 1280), rnd(1280, 9 * 1280, s=0.01), rnd(1280)
    return lambda: ops.conv3x3(x, w, bias=b)


def case_conv_h16():
    x, w, b = rnd(8, 16, 16, 1280), rnd(1280, 9 * 1280, s=0.01), rnd(1280)
    return lambda: ops.conv3x3(x, w, bias=b)


def case_conv_h64():
    x, w, b = rnd(8, 64, 64, 320), rnd(320, 9 * 320, s=0.02), rnd(320)
    return lambda: ops.conv3x3(x, w, bias=b)


def case_gn():
    x, g, b = rnd(8, 4096, 320), rnd(320), rnd(320)
    dy = rnd(8, 4096, 320)

    def run():
        y, st = ops.groupnorm(x, g, b, 32, 1e-5, True)
        ops.groupnorm_bwd(dy, x, g, b, st, 32, 1e-5, True, add=dy)
    return run


def case_gn_small():
    x, g, b = rnd(8, 64, 1280), rnd(1280), rnd(1280)
    dy = rnd(8, 64, 1280)

    def run():
        y, st = ops.groupnorm(x, g, b, 32, 1e-5, True)
        ops.groupnorm_bwd(dy, x, g, b, st, 32, 1e-5, True, add=dy)
    return run


def case_attn():
    B, H, N, d = 8, 8, 4096, 40
    Cc = H * d
    qkv = rnd(B, N, 3 * Cc)
    q, k, v = qkv[..., :Cc], qkv[..., Cc:2 * Cc], qkv[..., 2 * Cc:]
    do = rnd(B, N, Cc)

    def run():
        o, lse = ops.attn_fwd(q, k, v, H)
        ops.attn_bwd(q, k, v, o, do, lse, H)
    return run


def _attn_case(B, H, Nq, Nk, d, causal=False, dq_out=None):
    Cc = H * d
    if Nq == Nk:
        qkv = rnd(B, Nq, 3 * Cc)
        q, k, v = qkv[..., :Cc], qkv[..., Cc:2 * Cc], qkv[..., 2 * Cc:]
    else:
        q, kv = rnd(B, Nq, Cc), rnd(B, Nk, 2 * Cc)
        k, v = kv[..., :Cc], kv[..., Cc:]
    do = rnd(B, Nq, Cc)

    def run():
        o, lse = ops.attn_fwd(q, k, v, H, causal=causal)
        ops.attn_bwd(q, k, v, o, do, lse, H, causal=causal, dq_out=dq_out)
    return run


def case_attn_d80():    # 32x32 level self-attention
    return _attn_case(8, 8, 1024, 1024, 80)


def case_attn_d160():   # 16x16 level self-attention
    return _attn_case(8, 8, 256, 256, 160)


def case_attn_cross():  # 64x64 level cross-attention over the 77 text tokens: fp16 dQ straight out of the kernel
    return _attn_case(8, 8, 4096, 77, 40, dq_out=True)


def case_attn_cross_d80():
    return _attn_case(8, 8, 1024, 77, 80, dq_out=True)


def case_attn_clip():   # text encoder: causal, 12 heads x 64, 77 tokens
    return _attn_case(8, 12, 77, 77, 64, causal=True, dq_out=True)


def case_gn_mid():      # 32x32 level: the single-launch group-owner kernel
    x, g, b = rnd(8, 1024, 640), rnd(640), rnd(640)
    dy = rnd(8, 1024, 640)

    def run():
        y, st = ops.groupnorm(x, g, b, 32, 1e-5, True)
        ops.groupnorm_bwd(dy, x, g, b, st, 32, 1e-5, True, add=dy)
    return run


def case_clip_glue():   # text-encoder LayerNorm + LoRA kernels and the LoRA gradient kernel at 8 x 77 rows
    M, D, R, RP = 616, 768, 12, 16
    x, gamma, beta = rnd(M, D, dtype=torch.float32), rnd(D, dtype=torch.float32), rnd(D, dtype=torch.float32)
    A = rnd(R, D, dtype=torch.float32, s=0.3)
    y_ext, dy_ext, dqkv = rnd(M, D + RP), rnd(M, D + RP), rnd(M, 3 * D)
    g32, g16 = rnd(M, D, dtype=torch.float32), rnd(M, D)
    dB, dA = torch.zeros(3, D, 4, device=dev), torch.zeros(R, D, device=dev)

    def run():
        st = ops.layernorm_lora_fwd(x, gamma, beta, A, y_ext, RP)
        ops.layernorm_bwd_clip(dy_ext, x, gamma, st, add=g32, out=g32, out16=g16, lora_a=A)
        C.call("tb_lora_grad", C.ptr(dqkv), C.ptr(y_ext), C.ptr(dy_ext), D + RP, C.ptr(dB), C.ptr(dA), M, 3, 7, D, 4,
               1.0, C.stream_ptr())
    return run


def _clip_glue_parts():
    M, D, R, RP = 616, 768, 12, 16
    x, gamma, beta = rnd(M, D, dtype=torch.float32), rnd(D, dtype=torch.float32), rnd(D, dtype=torch.float32)
    A = rnd(R, D, dtype=torch.float32, s=0.3)
    y_ext, dy_ext, dqkv = rnd(M, D + RP), rnd(M, D + RP), rnd(M, 3 * D)
    g32, g16 = rnd(M, D, dtype=torch.float32), rnd(M, D)
    dB, dA = torch.zeros(3, D, 4, device=dev), torch.zeros(R, D, device=dev)
    st = ops.layernorm_lora_fwd(x, gamma, beta, A, y_ext, RP)
    return {
        "fwd": lambda: ops.layernorm_lora_fwd(x, gamma, beta, A, y_ext, RP),
        "bwd": lambda: ops.layernorm_bwd_clip(dy_ext, x, gamma, st, add=g32, out=g32, out16=g16, lora_a=A),
        "bwd_plain": lambda: ops.layernorm_bwd_clip(dy_ext[:, :D], x, gamma, st, add=g32, out=g32, out16=g16),
        "grad": lambda: C.call("tb_lora_grad", C.ptr(dqkv), C.ptr(y_ext), C.ptr(dy_ext), D + RP, C.ptr(dB), C.ptr(dA),
                               M, 3, 7, D, 4, 1.0, C.stream_ptr()),
    }


def case_clip_ln_fwd():
    return _clip_glue_parts()["fwd"]


def case_clip_ln_bwd():
    return _clip_glue_parts()["bwd"]


def case_clip_ln_bwd_plain():
    return _clip_glue_parts()["bwd_plain"]


def case_clip_lora_grad():
    return _clip_glue_parts()["grad"]


def case_ln():
    x, g, b = rnd(32768, 320), rnd(320), rnd(320)
    dy = rnd(32768, 320)

    def run():
        y, st = ops.layernorm(x, g, b)
        ops.layernorm_bwd(dy, x, g, st, add=dy)
    return run


def case_geglu():
    h, dg = rnd(32768, 2560), rnd(32768, 1280)

    def run():
        ops.geglu(h)
        ops.geglu_bwd(dg, h)
    return run


def case_unet_lora():   # --unet_params_to_train crossattn_kv: 32 rank-4 adapters on the fused K/V projection, SD-1.5
    M, ctx, r = 616, 768, 4
    widths = [320, 320, 640, 640, 1280, 1280, 1280] + [1280] * 3 + [640] * 3 + [320] * 3
    KV, n_ad = 2 * sum(widths), 2 * len(widths)
    blk, off = [], [0]
    for i, w in enumerate(widths):
        blk += [2 * i] * w + [2 * i + 1] * w
        off += [off[-1] + w, off[-1] + 2 * w]
    blk_t = torch.tensor(blk, dtype=torch.int32, device=dev)
    off_t = torch.tensor(off, dtype=torch.int32, device=dev)
    ehs, kv, dkv = rnd(M, ctx), rnd(M, KV), rnd(M, KV, s=0.01)
    A, Bm = rnd(n_ad * r, ctx, dtype=torch.float32, s=0.25), rnd(KV, r, dtype=torch.float32, s=0.02)
    dA, dB, d_ehs = torch.zeros_like(A), torch.zeros_like(Bm), torch.zeros(M, ctx, device=dev)

    def run():
        Z = ops.unet_lora_fwd(ehs, A, Bm, blk_t, kv, r, 1.0)
        ops.unet_lora_bwd(dkv, ehs, A, Bm, Z, blk_t, off_t, dA, dB, d_ehs, r, 1.0)
    return run


CASES = {k[5:]: v for k, v in globals().items() if k.startswith("case_")}
TIME = "--time" in sys.argv
if TIME:
    sys.argv.remove("--time")
want = sys.argv[1:] or list(CASES)
if TIME:  # plain CUDA-event timing (L2-warm, 20 launches back to back): quick A/B between builds
    for n in want:
        f = CASES[n]()
        for _ in range(3):
            f()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()  # 20 launches in one graph: no host launch latency in the number
        with torch.cuda.graph(g):
            for _ in range(20):
                f()
        g.replay()
        torch.cuda.synchronize()
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        print(f"{n:14s} {1e3 * e0.elapsed_time(e1) / 20:9.1f} us / call")
    sys.exit(0)
fns = [(n, CASES[n]()) for n in want]
for n, f in fns:
    f()
    f()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for n, f in fns:
    torch.cuda.nvtx.range_push(n)
    f()
    torch.cuda.nvtx.range_pop()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("cases:", [n for n, _ in fns])
