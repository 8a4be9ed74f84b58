"""How much of a step is fixed per-launch cost?  Chains of dependent launches replayed as one CUDA graph."""
import sys
import torch
sys.path.insert(0, ".")
from textboost_b200 import ops  # noqa: E402

dev = "cuda"


def chain(name, f, n=200):
    for _ in range(3):
        f()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            f()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:44s} {1e3 * e0.elapsed_time(e1) / n:7.2f} us / launch")


t = torch.randint(0, 1000, (8,), device=dev)
chain("timestep_embedding (1 tiny CTA)", lambda: ops.timestep_embedding(t, 320))
x = torch.randn(1024, device=dev, dtype=torch.float16)
chain("silu 1024 elements", lambda: ops.silu(x))
a = torch.randn(128, 64, device=dev, dtype=torch.float16)
w = torch.randn(128, 64, device=dev, dtype=torch.float16)
chain("gemm 128x128x64 (1 CTA, 1 k-block)", lambda: ops.gemm(a, w))
a2 = torch.randn(128 * 148, 64, device=dev, dtype=torch.float16)
chain("gemm 18944x128x64 (148 CTAs, 1 k-block)", lambda: ops.gemm(a2, w))
a3 = torch.randn(128 * 148, 640, device=dev, dtype=torch.float16)
w3 = torch.randn(128, 640, device=dev, dtype=torch.float16)
chain("gemm 18944x128x640 (148 CTAs, 10 k-blocks)", lambda: ops.gemm(a3, w3))
xx = torch.randn(8, 64, 1280, device=dev, dtype=torch.float16)
gm = torch.ones(1280, device=dev, dtype=torch.float16)
chain("groupnorm fwd 8x64x1280 (memset + 2 kernels)", lambda: ops.groupnorm(xx, gm, gm, 32, 1e-5, True), n=100)
