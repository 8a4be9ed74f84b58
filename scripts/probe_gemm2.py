"""GPU probe: timing of the UNet's GEMM shapes (B=8).  Diagnostic only."""
import sys
import torch
sys.path.insert(0, ".")
from textboost_b200 import ops  # noqa: E402

dev = "cuda"


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


shapes = [(32768, 2560, 320), (32768, 320, 1280), (32768, 960, 320), (32768, 320, 320), (8192, 5120, 640), (8192, 640, 2560),
          (8192, 1920, 640), (8192, 640, 640), (2048, 10240, 1280), (2048, 1280, 5120), (2048, 3840, 1280), (2048, 1280, 1280),
          (616, 2560, 768), (1232, 2304, 784), (1232, 3072, 768), (1232, 768, 3072)]
tot = 0.0
for (M, N, K) in shapes:
    a = torch.randn(M, K, device=dev, dtype=torch.float16)
    w = torch.randn(N, K, device=dev, dtype=torch.float16) / K ** 0.5
    bias = torch.randn(N, device=dev, dtype=torch.float16)
    res = torch.randn(M, N, device=dev, dtype=torch.float16)
    ms = time_it(lambda: ops.gemm(a, w, bias=bias))
    ms_r = time_it(lambda: ops.gemm(a, w, bias=bias, residual=res))
    mt = time_it(lambda: torch.nn.functional.linear(a, w, bias))
    tot += ms
    print(f"gemm M={M} N={N} K={K}: {ms:.3f} ms ({2.0 * M * N * K / ms / 1e9:.0f} TF/s)  +residual {ms_r:.3f} ms | cublas {mt:.3f} ms", flush=True)
print(f"sum {tot:.3f} ms")
