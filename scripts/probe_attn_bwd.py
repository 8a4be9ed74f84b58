"""A/B timings of the flash-attention backward at the UNet 64x64 self-attention shape (graph replay, GPU time)."""
import sys
import torch
sys.path.insert(0, ".")
from textboost_b200 import ops  # noqa: E402

B, H, N, d = 8, 8, 4096, 40
C = H * d
torch.manual_seed(0)
qkv = torch.randn(B, N, 3 * C, device="cuda", dtype=torch.float16)
q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
do = torch.randn(B, N, C, device="cuda", dtype=torch.float16)
o, lse = ops.attn_fwd(q, k, v, H)


def timeit(name, f, n=5):
    for _ in range(2):
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
    print(f"{name:28s} {1e3 * e0.elapsed_time(e1) / n:9.1f} us")


timeit("fwd", lambda: ops.attn_fwd(q, k, v, H))
timeit("bwd (dq)", lambda: ops.attn_bwd(q, k, v, o, do, lse, H))
timeit("bwd (no dq)", lambda: ops.attn_bwd(q, k, v, o, do, lse, H, need_dq=False))
for (n2, d2, h2) in ((1024, 80, 8), (256, 160, 8)):
    C2 = h2 * d2
    x = torch.randn(B, n2, 3 * C2, device="cuda", dtype=torch.float16)
    q2, k2, v2 = x[..., :C2], x[..., C2:2 * C2], x[..., 2 * C2:]
    do2 = torch.randn(B, n2, C2, device="cuda", dtype=torch.float16)
    o2, l2 = ops.attn_fwd(q2, k2, v2, h2)
    timeit(f"fwd N={n2} d={d2}", lambda: ops.attn_fwd(q2, k2, v2, h2))
    timeit(f"bwd N={n2} d={d2}", lambda: ops.attn_bwd(q2, k2, v2, o2, do2, l2, h2))
