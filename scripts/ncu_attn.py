"""one launch of the attention forward / backward at a UNet shape, for `ncu -k regex:attn_...` captures:
   python scripts/ncu_attn.py [d=40] [N=4096] [Nk=N] [fwd|bwd|both]"""
import sys
import torch
sys.path.insert(0, ".")
from textboost_b200 import ops  # noqa: E402

d = int(sys.argv[1]) if len(sys.argv) > 1 else 40
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
Nk = int(sys.argv[3]) if len(sys.argv) > 3 else N
what = sys.argv[4] if len(sys.argv) > 4 else "both"
B, H = 8, 8
C = H * d
torch.manual_seed(0)
q = torch.randn(B, N, C, device="cuda", dtype=torch.float16)
kv = torch.randn(B, Nk, 2 * C, device="cuda", dtype=torch.float16)
k, v = kv[..., :C], kv[..., C:]
do = torch.randn(B, N, C, device="cuda", dtype=torch.float16)
for _ in range(2):
    o, lse = ops.attn_fwd(q, k, v, H)
    if what != "fwd":
        ops.attn_bwd(q, k, v, o, do, lse, H)
torch.cuda.synchronize()
