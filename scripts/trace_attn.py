"""Debug: timeline of one forward-attention CTA (clock64 stamps) at the 64x64 self-attention shape.
v2 kernel events: 0/1 = MMA warp saw P_0 / P_1 of tile j; 3 = softmax warp starts waiting for S(j), 4 = S(j) seen,
5 = row maximum known (pass 1 done), 6 = last P chunk stored, 7 = P(j) published."""
import sys
import torch
sys.path.insert(0, ".")
from textboost_b200 import _cabi as C, ops  # noqa: E402

B, H, N, d = 8, 8, 4096, int(sys.argv[1]) if len(sys.argv) > 1 else 40
Cc = H * d
torch.manual_seed(0)
qkv = torch.randn(B, N, 3 * Cc, device="cuda", dtype=torch.float16)
q, k, v = qkv[..., :Cc], qkv[..., Cc:2 * Cc], qkv[..., 2 * Cc:]
for _ in range(2):
    ops.attn_fwd(q, k, v, H)
buf = torch.zeros(16 * 8 * 10, device="cuda", dtype=torch.int64)
C.call("tb_attn_debug_trace", C.ptr(buf))
ops.attn_fwd(q, k, v, H)
torch.cuda.synchronize()
C.call("tb_attn_debug_trace", None)
t = buf.view(16, 8, 10).cpu()
t0 = t[t > 0].min()
names = ["-", "-", "-", "sm:wait S", "sm:S seen", "sm:max done", "sm:P stored", "sm:P arrived"]
mma = ["top", "S0 free", "S0' issued", "-", "P0 seen", "-", "PV0 issued", "-"]
for j in range(5, 9):
    print(f"--- iteration {j}")
    print("  mma warp:", "  ".join(f"{mma[e]}:{int(t[j, e, 9] - t0)}" for e in range(8)))
    for e in (3, 4, 5, 6, 7):
        print(f"  {names[e]:14s}", " ".join(f"w{w}:{int(t[j, e, w] - t0):7d}" for w in range(8)))
per = (t[12, 7, 0] - t[4, 7, 0]).item() / 8
print(f"cycles per KV iteration (two 128x128 tiles), warp 0: {per:.0f}")
