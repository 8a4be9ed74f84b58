"""Debug: timeline of one backward-attention CTA (attn_bwd3_kernel, clock64 stamps) at the 64x64 self-attention shape."""
import sys
import torch
sys.path.insert(0, ".")
from textboost_b200 import _cabi as C, ops  # noqa: E402

B, H, N, d = 8, 8, 4096, int(sys.argv[1]) if len(sys.argv) > 1 else 40
Cc = H * d
torch.manual_seed(0)
qkv = torch.randn(B, N, 3 * Cc, device="cuda", dtype=torch.float16)
q, k, v = qkv[..., :Cc], qkv[..., Cc:2 * Cc], qkv[..., 2 * Cc:]
do = torch.randn(B, N, Cc, device="cuda", dtype=torch.float16)
o, lse = ops.attn_fwd(q, k, v, H)
for _ in range(2):
    ops.attn_bwd(q, k, v, o, do, lse, H)
buf = torch.zeros(16 * 8 * 20, device="cuda", dtype=torch.int64)
C.call("tb_attn_debug_trace", C.ptr(buf))
ops.attn_bwd(q, k, v, o, do, lse, H)
torch.cuda.synchronize()
C.call("tb_attn_debug_trace", None)
t = buf.view(16, 8, 20).cpu()
t0 = t[t > 0].min()
A = ["q_full", "S free", "S' issued", "dP free", "dP' issued"]
Bn = ["top", "dS seen", "dK,dQ issued", "P' seen", "dV' issued"]
Sn = ["-", "wait dP", "dP seen", "dS stored", "P' stored", "published"]
for j in range(5, 9):
    print(f"--- iteration {j}")
    print("  mma-A:", "  ".join(f"{A[e]}:{int(t[j, e, 17] - t0)}" for e in range(5)))
    print("  mma-B:", "  ".join(f"{Bn[e]}:{int(t[j, e, 18] - t0)}" for e in range(5)))
    for e in range(1, 6):
        print(f"  {Sn[e]:13s}", " ".join(f"w{w}:{int(t[j, e, w] - t0):6d}" for w in (0, 3, 5, 10, 15)))
print("cycles per iteration (warp 0):", (t[12, 5, 0] - t[4, 5, 0]).item() / 8)
