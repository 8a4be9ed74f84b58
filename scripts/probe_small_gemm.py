"""Graph-replay timings of the latency-bound GEMM shapes (text encoder at M = 616, deep UNet levels).
Run twice: plain, and with TB_GEMM_NO_SPLITK=1, to tune the split-K heuristic."""
import sys
import torch
sys.path.insert(0, ".")
from textboost_b200 import _cabi as C, ops  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
SHAPES = [(616, 2304, 784, False), (616, 768, 768, True), (616, 3072, 768, False), (616, 768, 3072, True),
          (616, 784, 2304, False), (616, 3072, 768, False), (1232, 768, 3072, True), (512, 1280, 1280, False),
          (512, 1280, 2560, False), (512, 1280, 5120, False), (512, 10240, 1280, False), (2048, 1280, 1280, False),
          (2048, 1280, 5120, False), (8192, 640, 640, False), (77 * 8, 2560, 768, False)]
for (M, N, K, f32) in SHAPES:
    a = torch.randn(M, K, device=dev, dtype=torch.float16)
    w = torch.randn(N, K, device=dev, dtype=torch.float16) * 0.03
    b = torch.randn(N, device=dev, dtype=torch.float16)
    r = torch.randn(M, N, device=dev, dtype=torch.float32 if f32 else torch.float16)
    kw = dict(bias=b, residual=r, out_kind=C.TB_OUT_F32 if f32 else C.TB_OUT_F16)
    f = lambda: ops.gemm(a, w, **kw)  # noqa: E731
    for _ in range(3):
        f()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            f()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / 20
    print(f"M={M:5d} N={N:5d} K={K:5d} f32={int(f32)}  {us:7.1f} us  {2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s")
