"""GPU probe: GroupNorm(+SiLU) forward / input-gradient at the UNet's shapes, graph replay of 20 calls each.
Environment knobs (csrc/norm.cu): TB_GN_FWD_VARIANT, TB_GN_BWD_VARIANT (U*10 + min blocks per SM), TB_GN_CTAS_PER_SM."""
import sys

import torch

sys.path.insert(0, ".")
from textboost_b200 import ops  # noqa: E402

dev = "cuda"
SHAPES = [(8, 4096, 320), (8, 1024, 640), (8, 256, 1280), (8, 64, 1280), (8, 4096, 640), (8, 4096, 960), (8, 1024, 1920),
          (8, 256, 2560)]
if "--big" in sys.argv:  # the shapes the cluster variant of the group-owner kernel covers
    SHAPES = [(8, 4096, 320), (8, 4096, 640), (8, 4096, 960), (8, 1024, 1920)]


def graph_us(fn, n=20):
    for _ in range(3):
        fn()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / n


tot_f = tot_b = 0.0
for B, HW, C in SHAPES:
    x = torch.randn(B, HW, C, device=dev, dtype=torch.float16)
    dy = torch.randn_like(x)
    add = torch.randn_like(x)
    gamma = torch.randn(C, device=dev, dtype=torch.float16)
    beta = torch.randn(C, device=dev, dtype=torch.float16)
    _, st = ops.groupnorm(x, gamma, beta, 32, 1e-5, True)
    f = graph_us(lambda: ops.groupnorm(x, gamma, beta, 32, 1e-5, True))
    b = graph_us(lambda: ops.groupnorm_bwd(dy, x, gamma, beta, st, 32, 1e-5, True, add=add))
    mb = B * HW * C * 2 / 1e6
    tot_f += f
    tot_b += b
    print(f"[{B},{HW},{C}] fwd {f:6.1f} us ({3 * mb / f / 1e3:5.2f} TB/s)   bwd {b:6.1f} us ({6 * mb / b / 1e3:5.2f} TB/s)")
print(f"sum fwd {tot_f:.1f} us  bwd {tot_b:.1f} us")
