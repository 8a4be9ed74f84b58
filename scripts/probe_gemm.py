"""GPU probe: tcgen05 GEMM / implicit-GEMM conv3x3 vs torch fp32, with timings.  Diagnostic only."""
import sys
import time

import torch

sys.path.insert(0, ".")
from textboost_b200 import _cabi as C  # noqa: E402
from textboost_b200 import ops  # noqa: E402

torch.manual_seed(0)
dev = "cuda"
print("device", torch.cuda.get_device_name(0), "lib version", C.lib().tb_version())


def relerr(a, b):
    return ((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-9)).item()


def time_it(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


ok = True
for (M, N, K) in [(128, 32, 64), (128, 64, 64), (128, 128, 64), (128, 256, 128), (256, 160, 320),
                  (616, 768, 768), (8, 1280, 320), (32768, 320, 320), (8192, 1280, 640),
                  (2048, 10240, 1280), (32768, 2560, 320), (616, 2304, 768), (1000, 136, 72)]:
    a = torch.randn(M, K, device=dev, dtype=torch.float16)
    w = torch.randn(N, K, device=dev, dtype=torch.float16) / K ** 0.5
    ref = a.float() @ w.float().t()
    out = ops.gemm(a, w)
    torch.cuda.synchronize()
    e = relerr(out, ref)
    ms = time_it(lambda: ops.gemm(a, w)) if M * N * K > 1e8 else 0.0
    tf = 2.0 * M * N * K / (ms * 1e-3) / 1e12 if ms else 0.0
    good = e < 2e-3
    ok &= good
    print(f"gemm M={M} N={N} K={K} relerr={e:.2e} {'OK' if good else 'FAIL'}  {ms:.3f} ms {tf:.1f} TF/s")
    if not good:
        d = (out.float() - ref).abs()
        idx = d.argmax().item()
        print("   worst at", idx // N, idx % N, out.flatten()[idx].item(), ref.flatten()[idx].item())
        bad_rows = (d.max(dim=1).values > 1e-2 * ref.abs().max()).nonzero().flatten()
        bad_cols = (d.max(dim=0).values > 1e-2 * ref.abs().max()).nonzero().flatten()
        print("   bad rows", bad_rows[:16].tolist(), len(bad_rows), "bad cols", bad_cols[:16].tolist(), len(bad_cols))

# epilogue
M, N, K = 1024, 320, 640
a = torch.randn(M, K, device=dev, dtype=torch.float16)
w = torch.randn(N, K, device=dev, dtype=torch.float16) / K ** 0.5
bias = torch.randn(N, device=dev, dtype=torch.float16)
rowvec = torch.randn(M // 256, N, device=dev, dtype=torch.float16)
res = torch.randn(M, N, device=dev, dtype=torch.float16)
ref = torch.nn.functional.silu(0.5 * (a.float() @ w.float().t()) + bias.float() +
                               rowvec.float().repeat_interleave(256, 0)) + res.float()
out = ops.gemm(a, w, bias=bias, rowvec=rowvec, rows_per_group=256, residual=res, alpha=0.5,
               act=C.TB_ACT_SILU)
e = relerr(out, ref)
ok &= e < 2e-3
print(f"gemm epilogue silu relerr={e:.2e}")
acc = torch.ones(M, N, device=dev, dtype=torch.float32)
ops.gemm(a, w, out=acc, out_kind=C.TB_OUT_F32_ACC)
e = relerr(acc, 1.0 + a.float() @ w.float().t())
ok &= e < 1e-3
print(f"gemm f32 accumulate relerr={e:.2e}")
for act, f in [(C.TB_ACT_QUICK_GELU, lambda x: x * torch.sigmoid(1.702 * x)),
               (C.TB_ACT_GELU, torch.nn.functional.gelu)]:
    out = ops.gemm(a, w, bias=bias, act=act)
    e = relerr(out, f(a.float() @ w.float().t() + bias.float()))
    ok &= e < 2e-3
    print(f"gemm act {act} relerr={e:.2e}")

# conv3x3
for (B, H, W, Cin, Cout) in [(1, 8, 8, 64, 64), (2, 8, 8, 128, 160), (2, 16, 16, 64, 128),
                             (1, 32, 32, 64, 320), (1, 64, 64, 64, 32), (3, 8, 8, 64, 64),
                             (8, 64, 64, 320, 320), (8, 32, 32, 640, 640), (8, 16, 16, 1280, 1280),
                             (8, 8, 8, 2560, 1280), (8, 64, 64, 960, 320)]:
    x = torch.randn(B, H, W, Cin, device=dev, dtype=torch.float16)
    wt = torch.randn(Cout, Cin, 3, 3, device=dev, dtype=torch.float16) / (9 * Cin) ** 0.5
    bias = torch.randn(Cout, device=dev, dtype=torch.float16)
    wk = wt.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
    ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).float(), wt.float(), bias.float(),
                                     padding=1).permute(0, 2, 3, 1)
    out = ops.conv3x3(x, wk, bias=bias)
    torch.cuda.synchronize()
    e = relerr(out, ref)
    flops = 2.0 * B * H * W * Cout * 9 * Cin
    ms = time_it(lambda: ops.conv3x3(x, wk, bias=bias)) if flops > 1e9 else 0.0
    tf = flops / (ms * 1e-3) / 1e12 if ms else 0.0
    good = e < 2e-3
    ok &= good
    print(f"conv3x3 B={B} H={H} W={W} Cin={Cin} Cout={Cout} relerr={e:.2e} {'OK' if good else 'FAIL'} {ms:.3f} ms {tf:.1f} TF/s")

print("ALL OK" if ok else "SOME FAILED")
