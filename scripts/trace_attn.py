"""Debug: timeline of one forward-attention CTA (clock64 stamps) at the 64x64 self-attention shape."""
import sys
import torch
sys.path.insert(0, ".")
from textboost_b200 import _cabi as C, ops  # noqa: E402

B, H, N, d = 8, 8, 4096, 40
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
names = ["mma:S issued", "mma:P seen", "mma:PV issued", "sm:wait S", "sm:S seen", "sm:S loaded", "sm:max xchg", "sm:P arrived"]
for j in range(2, 8):
    print(f"--- iteration {j}")
    for e in range(8):
        ws = [9] if e < 3 else [0, 3, 4, 7]
        print(f"  {names[e]:14s}", " ".join(f"w{w}:{int(t[j, e, w] - t0):7d}" for w in ws))
