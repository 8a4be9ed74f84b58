"""Single launches of the step's representative kernels at BASELINE shapes (SD-1.5, B=8), for

  ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/rNN_cases \
      python scripts/ncu_cases.py [case ...]

Each case runs twice untimed, then once between cudaProfilerStart/Stop.  Numbers printed under ncu are not
bench values.  Cases: gemm_k320 gemm_m2048 gemm_clip conv_h8 conv_h64 gn attn ln geglu
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
