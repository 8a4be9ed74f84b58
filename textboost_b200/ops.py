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
