"""One forward + one backward flash-attention launch at the UNet's 64x64 self-attention shape, for
ncu --set full -k regex:attn_(fwd|bwd)_kernel captures."""
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
for _ in range(2):
    o, lse = ops.attn_fwd(q, k, v, H)
    ops.attn_bwd(q, k, v, o, do, lse, H)
torch.cuda.synchronize()
