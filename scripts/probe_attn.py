"""GPU probe: flash attention fwd/bwd vs torch fp32 autograd.  Diagnostic only."""
import sys
import torch
sys.path.insert(0, ".")
from textboost_b200 import ops  # noqa: E402

torch.manual_seed(0)
dev = "cuda"


def relerr(a, b):
    return ((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-9)).item()


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


ok = True
cases = [(1, 1, 128, 128, 64), (1, 2, 128, 128, 40), (2, 2, 256, 256, 40), (1, 2, 256, 77, 40),
         (1, 2, 64, 64, 160), (2, 2, 256, 256, 160), (1, 2, 256, 77, 160), (2, 2, 256, 256, 80),
         (1, 2, 300, 200, 80), (2, 8, 1024, 1024, 80), (2, 8, 4096, 4096, 40), (8, 8, 4096, 77, 40),
         (8, 8, 4096, 4096, 40)]
for (B, H, Nq, Nk, d) in cases:
    C = H * d
    self_attn = Nq == Nk
    if self_attn:
        qkv = torch.randn(B, Nq, 3 * C, device=dev, dtype=torch.float16)
        q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    else:
        q = torch.randn(B, Nq, C, device=dev, dtype=torch.float16)
        kv = torch.randn(B, Nk, 2 * C, device=dev, dtype=torch.float16)
        k, v = kv[..., :C], kv[..., C:]
    do = torch.randn(B, Nq, C, device=dev, dtype=torch.float16)
    big = B * H * Nq * Nk > 5e8
    rdt = torch.float16 if big else torch.float32

    def heads(t):
        return t.reshape(B, -1, H, d).transpose(1, 2).to(rdt).detach().requires_grad_(True)
    qr, kr, vr = heads(q), heads(k), heads(v)
    oref = torch.nn.functional.scaled_dot_product_attention(qr, kr, vr)
    oref.backward(do.reshape(B, Nq, H, d).transpose(1, 2).to(rdt))
    oref2 = oref.transpose(1, 2).reshape(B, Nq, C)
    o, lse = ops.attn_fwd(q, k, v, H)
    torch.cuda.synchronize()
    e_o = relerr(o, oref2)
    dq, dk, dv = ops.attn_bwd(q, k, v, o, do, lse, H)
    torch.cuda.synchronize()
    e_dq = relerr(dq, qr.grad.transpose(1, 2).reshape(B, Nq, C))
    e_dk = relerr(dk, kr.grad.transpose(1, 2).reshape(B, Nk, C))
    e_dv = relerr(dv, vr.grad.transpose(1, 2).reshape(B, Nk, C))
    tol = 5e-3 if big else 3e-3
    good = max(e_o, e_dq, e_dk, e_dv) < tol
    ok &= good
    msg = f"attn B={B} H={H} Nq={Nq} Nk={Nk} d={d} o={e_o:.1e} dq={e_dq:.1e} dk={e_dk:.1e} dv={e_dv:.1e} {'OK' if good else 'FAIL'}"
    if big:
        tf = time_it(lambda: ops.attn_fwd(q, k, v, H))
        tb = time_it(lambda: ops.attn_bwd(q, k, v, o, do, lse, H))
        fl = 4.0 * B * H * Nq * Nk * d
        msg += f" fwd {tf:.3f} ms ({fl / tf / 1e9:.0f} TF/s) bwd {tb:.3f} ms ({2.5 * fl / tb / 1e9:.0f} TF/s)"
        qh, kh, vh = heads(q), heads(k), heads(v)
        tt = time_it(lambda: torch.nn.functional.scaled_dot_product_attention(qh, kh, vh))
        msg += f" | torch sdpa fwd {tt:.3f} ms"
    print(msg, flush=True)
print("ALL OK" if ok else "SOME FAILED")
