"""GPU probe: flash attention fwd/bwd timing at the UNet shapes (B=8).  Diagnostic only."""
import sys
import torch
sys.path.insert(0, ".")
from textboost_b200 import ops  # noqa: E402

torch.manual_seed(0)
dev = "cuda"


def time_it(fn, iters=10):
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


SHAPES = [(8, 8, 4096, 4096, 40), (8, 8, 4096, 77, 40), (8, 8, 1024, 1024, 80), (8, 8, 1024, 77, 80),
          (8, 8, 256, 256, 160), (8, 8, 256, 77, 160), (16, 12, 77, 77, 64), (4, 5, 9216, 9216, 64)]
if len(sys.argv) > 1 and sys.argv[1] == "self":  # the long self-attention shapes only
    SHAPES = [s for s in SHAPES if s[2] == s[3] and s[2] >= 1024]
for (B, H, Nq, Nk, d) in SHAPES:
    C = H * d
    if Nq == Nk:
        qkv = torch.randn(B, Nq, 3 * C, device=dev, dtype=torch.float16)
        q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    else:
        q = torch.randn(B, Nq, C, device=dev, dtype=torch.float16)
        kv = torch.randn(B, Nk, 2 * C, device=dev, dtype=torch.float16)
        k, v = kv[..., :C], kv[..., C:]
    do = torch.randn(B, Nq, C, device=dev, dtype=torch.float16)
    o, lse = ops.attn_fwd(q, k, v, H)
    if B * H * Nq * Nk <= 8 * 8 * 4096 * 4096:
        def hd(t):
            return t.reshape(B, -1, H, d).transpose(1, 2)
        oref = torch.nn.functional.scaled_dot_product_attention(hd(q).float(), hd(k).float(), hd(v).float())
        err = ((o.float().view(B, Nq, H, d).transpose(1, 2) - oref).abs().max() / oref.abs().max()).item()
        del oref
    else:
        err = float("nan")
    tf = time_it(lambda: ops.attn_fwd(q, k, v, H))
    tb = time_it(lambda: ops.attn_bwd(q, k, v, o, do, lse, H))
    fl = 4.0 * B * H * Nq * Nk * d

    def heads(t):
        return t.reshape(B, -1, H, d).transpose(1, 2).contiguous()
    qh, kh, vh = heads(q), heads(k), heads(v)
    tt = time_it(lambda: torch.nn.functional.scaled_dot_product_attention(qh, kh, vh))
    print(f"attn B={B} H={H} Nq={Nq} Nk={Nk} d={d}: fwd {tf:.3f} ms ({fl / tf / 1e9:.0f} TF/s) bwd {tb:.3f} ms "
          f"({2 * fl / tb / 1e9:.0f} TF/s alg) | torch sdpa fwd {tt:.3f} ms | fwd err {err:.1e}", flush=True)
