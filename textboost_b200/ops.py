"""Thin tensor-level wrappers over the C ABI (one Python function per entry point).

PyTorch is used for device memory and streams only; every function here allocates its output with
torch.empty and enqueues exactly the kernels of one C-ABI call on the current stream.
"""
from __future__ import annotations

import ctypes

import torch

from . import _cabi as C

F16 = torch.float16
F32 = torch.float32


def _epilogue(bias=None, rowvec=None, rows_per_group=1, residual=None, alpha=1.0,
              act=C.TB_ACT_NONE, out_kind=C.TB_OUT_F16):
    ep = C.Epilogue()
    ep.bias = bias.data_ptr() if bias is not None else None
    ep.rowvec = rowvec.data_ptr() if rowvec is not None else None
    ep.rows_per_group = int(rows_per_group)
    ep.residual = residual.data_ptr() if residual is not None else None
    ep.ldr = residual.stride(-2) if residual is not None else 0
    ep.alpha = float(alpha)
    ep.act = int(act)
    ep.out_kind = int(out_kind)
    return ep


def gemm(a: torch.Tensor, w: torch.Tensor, *, bias=None, rowvec=None, rows_per_group=1,
         residual=None, alpha=1.0, act=C.TB_ACT_NONE, out=None, out_kind=C.TB_OUT_F16):
    """out[M,N] = epilogue(a[M,K] @ w[N,K]^T).  a may be a 2-D view with row stride >= K."""
    assert a.dtype == F16 and w.dtype == F16 and a.dim() == 2 and w.dim() == 2
    assert a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K, (a.shape, w.shape)
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=F16 if out_kind == C.TB_OUT_F16 else F32)
    if residual is not None:
        assert residual.dtype == F16 and residual.stride(-1) == 1 and residual.shape == (M, N)
    ep = _epilogue(bias, rowvec, rows_per_group, residual, alpha, act, out_kind)
    C.call("tb_gemm_f16", C.ptr(a), a.stride(0), C.ptr(w), w.stride(0), C.ptr(out), out.stride(0),
           M, N, K, ctypes.byref(ep), C.stream_ptr())
    return out


def conv3x3(x: torch.Tensor, w: torch.Tensor, *, bias=None, rowvec=None, residual=None,
            act=C.TB_ACT_NONE, out=None):
    """x [B,H,W,Cin] fp16 contiguous NHWC, w [Cout, 9*Cin] (tap-major) -> [B,H,W,Cout]."""
    assert x.dtype == F16 and w.dtype == F16 and x.is_contiguous() and w.is_contiguous()
    B, H, W, Cin = x.shape
    Cout = w.shape[0]
    assert w.shape[1] == 9 * Cin
    if out is None:
        out = torch.empty((B, H, W, Cout), device=x.device, dtype=F16)
    res2d = residual.view(-1, Cout) if residual is not None else None
    ep = _epilogue(bias, rowvec, H * W, res2d, 1.0, act, C.TB_OUT_F16)
    C.call("tb_conv3x3_f16", C.ptr(x), C.ptr(w), C.ptr(out), B, H, W, Cin, Cout, ctypes.byref(ep),
           C.stream_ptr())
    return out


def attn_fwd(q, k, v, heads, scale=None, out=None):
    """q [B,Nq,*] k,v [B,Nk,*] fp16 views whose last dim holds heads*d columns (row stride free).
    Returns (o [B,Nq,heads*d], lse [B,heads,Nq])."""
    B, Nq, Ch = q.shape
    Nk = k.shape[1]
    d = Ch // heads
    scale = d ** -0.5 if scale is None else scale
    assert q.stride(2) == 1 and k.stride(2) == 1 and v.stride(2) == 1
    assert q.stride(0) == Nq * q.stride(1) and k.stride(0) == Nk * k.stride(1) and v.stride(0) == Nk * v.stride(1)
    if out is None:
        out = torch.empty((B, Nq, Ch), device=q.device, dtype=F16)
    lse = torch.empty((B, heads, Nq), device=q.device, dtype=F32)
    C.call("tb_attn_fwd_f16", C.ptr(q), q.stride(1), C.ptr(k), k.stride(1), C.ptr(v), v.stride(1),
           C.ptr(out), out.stride(1), C.ptr(lse), B, heads, Nq, Nk, d, scale, C.stream_ptr())
    return out, lse


def attn_bwd(q, k, v, o, do, lse, heads, scale=None, need_dq=True, dk=None, dv=None):
    """Returns (dq_acc fp32 [B,Nq,C] or None, dk, dv fp16 [B,Nk,C])."""
    B, Nq, Ch = q.shape
    Nk = k.shape[1]
    d = Ch // heads
    scale = d ** -0.5 if scale is None else scale
    delta = torch.empty((B, heads, Nq), device=q.device, dtype=F32)
    dq = torch.empty((B, Nq, Ch), device=q.device, dtype=F32) if need_dq else None
    if dk is None:
        dk = torch.empty((B, Nk, Ch), device=q.device, dtype=F16)
    if dv is None:
        dv = torch.empty((B, Nk, Ch), device=q.device, dtype=F16)
    assert do.stride(2) == 1 and o.stride(2) == 1
    C.call("tb_attn_bwd_f16", C.ptr(q), q.stride(1), C.ptr(k), k.stride(1), C.ptr(v), v.stride(1),
           C.ptr(o), o.stride(1), C.ptr(do), do.stride(1), C.ptr(lse), C.ptr(delta), C.ptr(dq),
           dq.stride(1) if dq is not None else 0, C.ptr(dk), dk.stride(1),
           C.ptr(dv), dv.stride(1), B, heads, Nq, Nk, d, scale, C.stream_ptr())
    return dq, dk, dv
