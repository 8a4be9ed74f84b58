"""TEST INFRASTRUCTURE ONLY: lets the ENGINES (ClipEngine, UNetEngine, TextBoostTrainer, FusedAdamW) run on CPU tensors
so that the CPU suite checks their orchestration — layer order, saved-activation bookkeeping, the hand-derived
activation-backward chain, LoRA packing, the optimiser tail — against the oracle step:

  * every SIMT entry point runs the PRODUCT source through the host build of the C ABI (kernel_host_emulation);
  * the four tensor-core wrappers (ops.gemm / conv3x3 / attn_fwd / attn_bwd: tcgen05 kernels, hardware only) are
    replaced by plain-PyTorch statements of their contracts, with the epilogue order of csrc/gemm.cu
    (alpha -> + bias + row vector -> activation -> + residual) and fp16 / fp32 output kinds;
  * the CUDA stream calls of the trainer become no-ops.
The tensor-core kernels themselves are checked on the GPU (tests/test_gpu_kernels.py, test_gpu_step.py)."""
import contextlib

import torch
import torch.nn.functional as F

from textboost_b200 import _cabi as C

F32 = torch.float32


def _h16():
    """The policy's 16-bit dtype, looked up at call time (fp16; bf16 when a test switched the policy)."""
    from textboost_b200.precision import POLICY
    return POLICY.act


def _act(y, act):
    if act == C.TB_ACT_SILU:
        return F.silu(y)
    if act == C.TB_ACT_QUICK_GELU:
        return y * torch.sigmoid(1.702 * y)
    if act == C.TB_ACT_GELU:
        return F.gelu(y)
    return y


def _finish(y, bias, rowvec, rows_per_group, residual, alpha, act, out, out_kind):
    y = y * alpha
    if bias is not None:
        y = y + bias.float()
    if rowvec is not None:
        y = y + rowvec.float().repeat_interleave(rows_per_group, dim=0)
    y = _act(y, act)
    if residual is not None:
        y = y + residual.float().reshape(y.shape)
    if out_kind == C.TB_OUT_F32_ACC:
        out.add_(y.reshape(out.shape))
        return out
    y = y.to(_h16() if out_kind == C.TB_OUT_F16 else F32)
    if out is not None:
        out.copy_(y.reshape(out.shape))
        return out
    return y


def gemm(a, w, *, bias=None, rowvec=None, rows_per_group=1, residual=None, alpha=1.0, act=C.TB_ACT_NONE, out=None,
         out_kind=C.TB_OUT_F16):
    assert a.dtype == _h16() and w.dtype == _h16() and a.stride(1) == 1 and w.stride(1) == 1 and w.shape[1] == a.shape[1]
    return _finish(a.float() @ w.float().t(), bias, rowvec, rows_per_group, residual, alpha, act, out, out_kind)


def conv3x3(x, w, *, bias=None, rowvec=None, residual=None, act=C.TB_ACT_NONE, out=None):
    B, H, W, Cin = x.shape
    Cout = w.shape[0]
    assert x.dtype == _h16() and x.is_contiguous() and w.shape[1] == 9 * Cin and Cin % 64 == 0
    wt = w.float().view(Cout, 3, 3, Cin).permute(0, 3, 1, 2)
    y = F.conv2d(x.float().permute(0, 3, 1, 2), wt, padding=1).permute(0, 2, 3, 1).reshape(B * H * W, Cout)
    res = residual.reshape(B * H * W, Cout) if residual is not None else None
    y = _finish(y, bias, rowvec, H * W, res, 1.0, act, None, C.TB_OUT_F16).view(B, H, W, Cout)
    if out is not None:
        out.copy_(y)
        return out
    return y


def _heads(t, heads):
    B, N, Ch = t.shape
    return t.float().view(B, N, heads, Ch // heads).transpose(1, 2)


def attn_fwd(q, k, v, heads, scale=None, out=None, causal=False):
    B, Nq, Ch = q.shape
    d = Ch // heads
    scale = d ** -0.5 if scale is None else scale
    s = _heads(q, heads) @ _heads(k, heads).transpose(-1, -2) * scale
    if causal:
        s = s.masked_fill(torch.ones(Nq, k.shape[1], dtype=torch.bool).triu(1), float("-inf"))
    lse = torch.logsumexp(s, -1) / 0.6931471805599453  # the kernels keep it in the log2 domain
    o = (torch.softmax(s, -1) @ _heads(v, heads)).transpose(1, 2).reshape(B, Nq, Ch).to(_h16())
    if out is not None:
        out.copy_(o)
        o = out
    return o, lse.contiguous()


def attn_bwd(q, k, v, o, do, lse, heads, scale=None, need_dq=True, dk=None, dv=None, causal=False, dq_out=None):
    B, Nq, Ch = q.shape
    d = Ch // heads
    scale = d ** -0.5 if scale is None else scale
    with torch.enable_grad():
        qf, kf, vf = (t.float().detach().clone().requires_grad_(True) for t in (q, k, v))
        s = _heads(qf, heads) @ _heads(kf, heads).transpose(-1, -2) * scale
        if causal:
            s = s.masked_fill(torch.ones(Nq, k.shape[1], dtype=torch.bool).triu(1), float("-inf"))
        out = (torch.softmax(s, -1) @ _heads(vf, heads)).transpose(1, 2).reshape(B, Nq, Ch)
        out.backward(do.float())
    dq = qf.grad.contiguous() if need_dq else None
    if dq_out is not None and dq_out is not False:  # the single-KV-tile contract: fp16 dQ written once, in place
        assert need_dq and k.shape[1] <= 128
        if dq_out is True:
            dq = dq.to(_h16())
        else:
            dq_out.copy_(dq)
            dq = dq_out
    gk, gv = kf.grad.to(_h16()), vf.grad.to(_h16())
    if dk is not None:
        dk.copy_(gk)
        gk = dk
    if dv is not None:
        dv.copy_(gv)
        gv = dv
    return dq, gk, gv


class _NoEvent:
    def record(self, stream=None):
        pass


class _NoStream:
    def wait_stream(self, other):
        pass

    def wait_event(self, event):
        pass

    def synchronize(self):
        pass


def install(monkeypatch, bf16: bool = False):
    """bf16=True: the bf16 host build of the SIMT sources and the policy switched to bf16 (K.install_abi_bf16)."""
    import kernel_host_emulation as K
    from textboost_b200 import ops
    (K.install_abi_bf16 if bf16 else K.install_abi)(monkeypatch)
    for name in ("gemm", "conv3x3", "attn_fwd", "attn_bwd"):
        monkeypatch.setattr(ops, name, globals()[name])
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: _NoStream())
    monkeypatch.setattr(torch.cuda, "Stream", lambda *a, **k: _NoStream())
    monkeypatch.setattr(torch.cuda, "Event", lambda *a, **k: _NoEvent())
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
