"""GPU parity tests, kernel level: every C-ABI entry point on the hot path against a plain PyTorch fp32
statement of the same op (floating-point kernels; tolerance stated per test).  All calls go through
libtextboost_b200.so via textboost_b200.ops (ctypes)."""
import math

import pytest
import torch
import torch.nn.functional as F

from conftest import act_dtype, tol_scale
from textboost_b200 import _cabi as C

pytestmark = pytest.mark.gpu

dev = "cuda"
F16 = act_dtype()  # fp16; bf16 when the file is re-run under the bf16 policy (tests/test_gpu_bf16_policy.py)
TOLX = tol_scale()  # 1 for fp16, 8 for bf16: relerr() below reports errors in units of the fp16 bounds written here


@pytest.fixture(scope="module", autouse=True)
def _need_gpu(built_lib):
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    from textboost_b200 import _cabi
    _cabi.call("tb_check_device")  # fails loudly if the library cannot run here


def relerr(a, b):
    return ((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-9)).item() / TOLX


# fp16 inputs, fp32 accumulation, fp16 output: one rounding of the result => 2^-11 ~ 4.9e-4 of the max
TOL_GEMM = 2e-3


@pytest.mark.parametrize("M,N,K", [(128, 32, 64), (128, 256, 128), (256, 160, 320), (616, 768, 784),
                                   (8, 1280, 320), (32768, 320, 320), (2048, 10240, 1280), (1000, 136, 72),
                                   (616, 2304, 768), (77, 1280, 768)])
def test_gemm(M, N, K):
    from textboost_b200 import ops
    g = torch.Generator(device=dev).manual_seed(M + N + K)
    a = torch.randn(M, K, device=dev, dtype=F16, generator=g)
    w = torch.randn(N, K, device=dev, dtype=F16, generator=g) / K ** 0.5
    out = ops.gemm(a, w)
    assert relerr(out, a.float() @ w.float().t()) < TOL_GEMM


def test_gemm_epilogues_and_strides():
    from textboost_b200 import _cabi as C, ops
    torch.manual_seed(0)
    M, N, K = 1024, 320, 640
    big = torch.randn(M, K + 64, device=dev, dtype=F16)
    a = big[:, 32:32 + K]  # strided A view (row stride > K), 64-byte aligned start
    w = torch.randn(N, K, device=dev, dtype=F16) / K ** 0.5
    bias = torch.randn(N, device=dev, dtype=F16)
    rowvec = torch.randn(M // 256, N, device=dev, dtype=F16)
    res = torch.randn(M, N, device=dev, dtype=F16)
    ref = F.silu(0.5 * (a.float() @ w.float().t()) + bias.float() + rowvec.float().repeat_interleave(256, 0)) + res.float()
    out = ops.gemm(a, w, bias=bias, rowvec=rowvec, rows_per_group=256, residual=res, alpha=0.5, act=C.TB_ACT_SILU)
    assert relerr(out, ref) < TOL_GEMM
    acc = torch.ones(M, N, device=dev, dtype=torch.float32)
    ops.gemm(a, w, out=acc, out_kind=C.TB_OUT_F32_ACC)
    assert relerr(acc, 1.0 + a.float() @ w.float().t()) < 1e-5
    res32 = torch.randn(M, N, device=dev, dtype=torch.float32)
    o32 = ops.gemm(a, w, bias=bias, residual=res32, out_kind=C.TB_OUT_F32)
    assert relerr(o32, a.float() @ w.float().t() + bias.float() + res32) < 1e-5
    for act, f in [(C.TB_ACT_QUICK_GELU, lambda x: x * torch.sigmoid(1.702 * x)), (C.TB_ACT_GELU, F.gelu)]:
        out = ops.gemm(a, w, bias=bias, act=act)
        assert relerr(out, f(a.float() @ w.float().t() + bias.float())) < TOL_GEMM


@pytest.mark.parametrize("M,N,K", [(616, 768, 3072), (512, 1280, 10240), (77, 2560, 8192), (8, 1280, 8256),
                                   (1232, 784, 2304), (300, 160, 16384)])
def test_gemm_split_k(M, N, K):
    """Under-filled problems with a very long K loop (>= 128 k-blocks; shorter ones are left alone because the
    fix-up costs more than it saves) take the split-K path (per-stream workspace registered by the binding): same result as the unsplit kernel up to fp32 summation order, identical between replays (the
    partials are added in split order), every fused epilogue still applied by the fixing CTA."""
    import os
    from textboost_b200 import _cabi as C, ops
    g = torch.Generator(device=dev).manual_seed(M * 7 + N + K)
    a = torch.randn(M, K, device=dev, dtype=F16, generator=g)
    w = torch.randn(N, K, device=dev, dtype=F16, generator=g) / K ** 0.5
    bias = torch.randn(N, device=dev, dtype=F16, generator=g)
    res = torch.randn(M, N, device=dev, dtype=torch.float32, generator=g)
    ref = F.gelu(a.float() @ w.float().t() + bias.float()) + res
    out = ops.gemm(a, w, bias=bias, residual=res, act=C.TB_ACT_GELU, out_kind=C.TB_OUT_F32)
    assert relerr(out, ref) < 1e-3
    out2 = ops.gemm(a, w, bias=bias, residual=res, act=C.TB_ACT_GELU, out_kind=C.TB_OUT_F32)
    assert torch.equal(out, out2)
    h = ops.gemm(a, w, bias=bias)  # fp16 output through the TMA-store boxes
    assert relerr(h, a.float() @ w.float().t() + bias.float()) < TOL_GEMM
    acc = torch.full((M, N), 2.0, device=dev, dtype=torch.float32)
    ops.gemm(a, w, out=acc, out_kind=C.TB_OUT_F32_ACC)
    assert relerr(acc, a.float() @ w.float().t() + 2.0) < 1e-3
    # without a workspace the same call runs unsplit: results agree to fp32 reassociation
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        C.stream_ptr()  # registers a workspace for s ...
        C.call("tb_set_workspace", C.c_void_p(s.cuda_stream), C.c_void_p(0), 0)  # ... and drops it again
        C._workspaces[(s.device.index, s.cuda_stream)] = None
        plain = ops.gemm(a, w, bias=bias, residual=res, act=C.TB_ACT_GELU, out_kind=C.TB_OUT_F32)
    s.synchronize()
    assert relerr(plain, out) < 1e-4


def test_gemm_rejects_bad_arguments():
    from textboost_b200 import ops
    a = torch.randn(16, 20, device=dev, dtype=F16)  # K % 8 != 0
    w = torch.randn(8, 20, device=dev, dtype=F16)
    with pytest.raises(RuntimeError, match="tb_gemm_f16"):
        ops.gemm(a, w)


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(1, 8, 8, 64, 64), (2, 8, 8, 128, 160), (2, 16, 16, 64, 128),
                                            (1, 32, 32, 64, 320), (1, 64, 64, 64, 32), (3, 8, 8, 64, 64),
                                            (2, 64, 64, 320, 320), (8, 8, 8, 2560, 1280), (2, 64, 64, 960, 320)])
def test_conv3x3(B, H, W, Cin, Cout):
    from textboost_b200 import ops
    from textboost_b200.unet import _conv_dgrad_weight, _conv_fwd_weight
    g = torch.Generator(device=dev).manual_seed(B * H + Cin)
    x = torch.randn(B, H, W, Cin, device=dev, dtype=F16, generator=g)
    wt = torch.randn(Cout, Cin, 3, 3, device=dev, dtype=F16, generator=g) / (9 * Cin) ** 0.5
    bias = torch.randn(Cout, device=dev, dtype=F16, generator=g)
    xr = x.permute(0, 3, 1, 2).float().requires_grad_(True)
    ref = F.conv2d(xr, wt.float(), bias.float(), padding=1)
    out = ops.conv3x3(x, _conv_fwd_weight(wt), bias=bias)
    assert relerr(out, ref.permute(0, 2, 3, 1)) < TOL_GEMM
    if Cout % 64 == 0:  # input gradient = the same kernel on the flipped / transposed weight
        dy = torch.randn(B, H, W, Cout, device=dev, dtype=F16, generator=g)
        ref.backward(dy.permute(0, 3, 1, 2).float())
        dx = ops.conv3x3(dy, _conv_dgrad_weight(wt))
        assert relerr(dx, xr.grad.permute(0, 2, 3, 1)) < TOL_GEMM


@pytest.mark.parametrize("B,H,Cin,Cout", [(2, 96, 64, 64), (1, 48, 128, 160), (3, 24, 64, 128), (2, 12, 128, 64),
                                          (1, 12, 1280, 1280), (1, 8, 64, 64), (3, 8, 64, 64), (1, 6, 64, 64)])
def test_conv3x3_non_power_of_two_planes(B, H, Cin, Cout):
    """768^2 images (configs[4]): 96 / 48 / 24 / 12-wide planes have no 128-pixel box; the implicit GEMM then runs
    96- (72-) pixel tiles inside the 128-row MMA.  Also batch sizes that do not fill a tile at the small levels.
    Bias + time-embedding row vector + fp16 residual as in ResnetBlock2D, forward and input gradient."""
    from textboost_b200 import ops
    from textboost_b200.unet import _conv_dgrad_weight, _conv_fwd_weight
    g = torch.Generator(device=dev).manual_seed(B * 1000 + H * 10 + Cin)
    x = torch.randn(B, H, H, Cin, device=dev, dtype=F16, generator=g)
    wt = torch.randn(Cout, Cin, 3, 3, device=dev, dtype=F16, generator=g) / (9 * Cin) ** 0.5
    bias = torch.randn(Cout, device=dev, dtype=F16, generator=g)
    temb = torch.randn(B, Cout, device=dev, dtype=F16, generator=g)
    res = torch.randn(B, H, H, Cout, device=dev, dtype=F16, generator=g)
    xr = x.permute(0, 3, 1, 2).float().requires_grad_(True)
    ref = F.conv2d(xr, wt.float(), bias.float(), padding=1) + temb.float()[:, :, None, None] + res.permute(0, 3, 1, 2).float()
    out = ops.conv3x3(x, _conv_fwd_weight(wt), bias=bias, rowvec=temb, residual=res)
    assert relerr(out, ref.permute(0, 2, 3, 1)) < TOL_GEMM
    if Cout % 64 == 0:
        dy = torch.randn(B, H, H, Cout, device=dev, dtype=F16, generator=g)
        ref.backward(dy.permute(0, 3, 1, 2).float())
        dx = ops.conv3x3(dy, _conv_dgrad_weight(wt))
        assert relerr(dx, xr.grad.permute(0, 2, 3, 1)) < TOL_GEMM


def test_conv3x3_split_k_deep_levels():
    """The 8x8 / 16x16 UNet levels (M = 512 / 2048 pixels, K = 9*1280): split along K."""
    from textboost_b200 import ops
    torch.manual_seed(3)
    for (B, H, Cin, Cout) in ((8, 8, 1280, 1280), (2, 16, 640, 320), (8, 8, 2560, 1280)):
        x = torch.randn(B, H, H, Cin, device=dev, dtype=F16)
        w4 = torch.randn(Cout, Cin, 3, 3, device=dev, dtype=F16) / (9 * Cin) ** 0.5
        bias = torch.randn(Cout, device=dev, dtype=F16)
        res = torch.randn(B, H, H, Cout, device=dev, dtype=F16)
        wk = w4.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous()
        y = ops.conv3x3(x, wk, bias=bias, residual=res)
        ref = F.conv2d(x.permute(0, 3, 1, 2).float(), w4.float(), bias.float(), padding=1).permute(0, 2, 3, 1) + res.float()
        assert relerr(y, ref) < TOL_GEMM
        assert torch.equal(y, ops.conv3x3(x, wk, bias=bias, residual=res))


def test_strided_conv_and_upsample_paths():
    """Downsample2D (im2col + GEMM / zero-stuff + conv dgrad) and Upsample2D (nearest x2 and its adjoint)."""
    from textboost_b200 import ops
    from textboost_b200.unet import _Conv3
    torch.manual_seed(1)
    B, H, Cc = 2, 16, 64
    x = torch.randn(B, H, H, Cc, device=dev, dtype=F16)
    wt = torch.randn(Cc, Cc, 3, 3, device=dev, dtype=F16) / (9 * Cc) ** 0.5
    bias = torch.randn(Cc, device=dev, dtype=F16)
    conv = _Conv3(wt, bias)
    xr = x.permute(0, 3, 1, 2).float().requires_grad_(True)
    ref = F.conv2d(xr, wt.float(), bias.float(), stride=2, padding=1)
    out = ops.gemm(ops.im2col3x3s2(x), conv.wk, bias=bias).view(B, H // 2, H // 2, Cc)
    assert relerr(out, ref.permute(0, 2, 3, 1)) < TOL_GEMM
    dy = torch.randn(B, H // 2, H // 2, Cc, device=dev, dtype=F16)
    ref.backward(dy.permute(0, 3, 1, 2).float())
    dx = conv.dgrad(ops.zero_stuff2x(dy))
    assert relerr(dx, xr.grad.permute(0, 2, 3, 1)) < TOL_GEMM
    up = ops.upsample2x(x)
    assert torch.equal(up, x.repeat_interleave(2, 1).repeat_interleave(2, 2))
    dup = torch.randn(B, 2 * H, 2 * H, Cc, device=dev, dtype=F16)
    ref_d = dup.float().view(B, H, 2, H, 2, Cc).sum((2, 4))
    assert relerr(ops.upsample2x_bwd(dup), ref_d) < 1e-3


# attention: P is rounded to fp16 before P V (as torch SDPA's flash kernels do); outputs fp16
TOL_ATTN = 3e-3


@pytest.mark.parametrize("B,H,Nq,Nk,d", [(1, 1, 128, 128, 64), (1, 2, 128, 128, 40), (2, 2, 256, 256, 40),
                                         (1, 2, 256, 77, 40), (1, 2, 64, 64, 160), (2, 2, 256, 256, 160),
                                         (1, 2, 256, 77, 160), (1, 2, 300, 200, 80), (2, 8, 1024, 1024, 80),
                                         (1, 8, 4096, 4096, 40), (2, 8, 4096, 77, 40), (1, 5, 2304, 2304, 64),
                                         (1, 4, 1024, 77, 160), (2, 8, 1024, 77, 80), (8, 8, 4096, 77, 40),
                                         # the two-query-tile forward kernel: a lone first tile (Nq = 384), padding
                                         # keys in the last KV tile, head_dim 64 / 80 / 96, 2 and 3 KV stages
                                         (1, 2, 384, 300, 40), (1, 2, 576, 576, 64), (1, 2, 300, 333, 80),
                                         (2, 1, 512, 1000, 96), (1, 3, 1280, 256, 48)])
def test_attention_fwd_bwd(B, H, Nq, Nk, d):
    from textboost_b200 import ops
    g = torch.Generator(device=dev).manual_seed(Nq + Nk + d)
    Cc = H * d
    if Nq == Nk:
        qkv = torch.randn(B, Nq, 3 * Cc, device=dev, dtype=F16, generator=g)
        q, k, v = qkv[..., :Cc], qkv[..., Cc:2 * Cc], qkv[..., 2 * Cc:]
    else:
        q = torch.randn(B, Nq, Cc, device=dev, dtype=F16, generator=g)
        kv = torch.randn(B, Nk, 2 * Cc, device=dev, dtype=F16, generator=g)
        k, v = kv[..., :Cc], kv[..., Cc:]
    do = torch.randn(B, Nq, Cc, device=dev, dtype=F16, generator=g)

    def heads(t):
        return t.reshape(B, -1, H, d).transpose(1, 2).float().detach().requires_grad_(True)

    qr, kr, vr = heads(q), heads(k), heads(v)
    s = (qr @ kr.transpose(-1, -2)) * d ** -0.5
    oref = torch.softmax(s, -1) @ vr
    oref.backward(do.reshape(B, Nq, H, d).transpose(1, 2).float())
    o, lse = ops.attn_fwd(q, k, v, H)
    assert relerr(o, oref.transpose(1, 2).reshape(B, Nq, Cc)) < TOL_ATTN
    lse_ref = torch.logsumexp(s.detach(), -1) * math.log2(math.e)
    assert (lse - lse_ref).abs().max().item() < 2e-2
    dq, dk, dv = ops.attn_bwd(q, k, v, o, do, lse, H)
    assert relerr(dq, qr.grad.transpose(1, 2).reshape(B, Nq, Cc)) < TOL_ATTN
    assert relerr(dk, kr.grad.transpose(1, 2).reshape(B, Nk, Cc)) < TOL_ATTN
    assert relerr(dv, vr.grad.transpose(1, 2).reshape(B, Nk, Cc)) < TOL_ATTN
    dq2, dk2, dv2 = ops.attn_bwd(q, k, v, o, do, lse, H, need_dq=False)  # first cross-attention: dK/dV only
    # (under-filled grids split the Q loop over CTAs that meet in fp32 red.adds: equal up to summation order)
    assert dq2 is None and relerr(dk2, dk) < 1e-3 and relerr(dv2, dv) < 1e-3
    if Nk <= 128:
        # a single KV tile: dQ stored once as fp16, here into a strided view (as the fused gradient tensor is)
        buf = torch.full((B, Nq, Cc + 8), 7.0, device=dev, dtype=F16)
        dq3, dk3, dv3 = ops.attn_bwd(q, k, v, o, do, lse, H, dq_out=buf[..., :Cc])
        assert dq3.dtype == F16 and dq3.data_ptr() == buf.data_ptr() and bool((buf[..., Cc:] == 7.0).all())
        assert relerr(dq3, qr.grad.transpose(1, 2).reshape(B, Nq, Cc)) < TOL_ATTN
        assert relerr(dq3, dq) < 2e-3 and relerr(dk3, dk) < 1e-3 and relerr(dv3, dv) < 1e-3
    else:
        with pytest.raises(RuntimeError, match="dQ16 needs Nk <= 128"):
            C.call("tb_attn_bwd_f16", C.ptr(q), q.stride(1), C.ptr(k), k.stride(1), C.ptr(v), v.stride(1), C.ptr(o),
                   o.stride(1), C.ptr(do), do.stride(1), C.ptr(lse), C.ptr(torch.empty(B, H, Nq, device=dev)), None, 0,
                   C.ptr(torch.empty(B, Nq, Cc, device=dev, dtype=F16)), Cc, C.ptr(dk), dk.stride(1), C.ptr(dv),
                   dv.stride(1), B, H, Nq, Nk, d, d ** -0.5, 0, C.stream_ptr())


@pytest.mark.parametrize("d", [40, 80])
def test_attention_fwd_growing_row_maximum(d):
    """The forward kernels keep a stale reference maximum until the row maximum grows by more than 2^8 and then
    rescale O and the row sum in TMEM: later KV tiles carry much larger scores here, so every row rescales several
    times; a row of all-equal scores and a constant V check the row sum itself."""
    from textboost_b200 import ops
    g = torch.Generator(device=dev).manual_seed(d)
    B, H, N = 1, 2, 1024
    Cc = H * d
    q = torch.randn(B, N, Cc, device=dev, dtype=F16, generator=g)
    k = torch.randn(B, N, Cc, device=dev, dtype=F16, generator=g)
    v = torch.randn(B, N, Cc, device=dev, dtype=F16, generator=g)
    k[:, 256:512] *= 4.0
    k[:, 640:] *= 12.0   # scores up to ~ +-80 in the last tiles
    q[:, 7] = 0          # an all-equal row: uniform softmax
    o, lse = ops.attn_fwd(q, k, v, H)

    def heads(t):
        return t.reshape(B, -1, H, d).transpose(1, 2).float()
    s = (heads(q) @ heads(k).transpose(-1, -2)) * d ** -0.5
    oref = (torch.softmax(s, -1) @ heads(v)).transpose(1, 2).reshape(B, N, Cc)
    assert relerr(o, oref) < TOL_ATTN
    assert (lse - torch.logsumexp(s, -1) * math.log2(math.e)).abs().max().item() < 2e-2
    assert (o[:, 7].float() - v.float().view(B, N, H, d).mean(1).reshape(B, Cc)).abs().max().item() < 2e-3


@pytest.mark.parametrize("B,H,N,d", [(2, 12, 77, 64), (3, 2, 128, 64), (1, 2, 300, 40), (16, 16, 77, 64)])
def test_attention_causal(B, H, N, d):
    """causal=1: the CLIP text-encoder attention (77 tokens, head_dim 64) on the same flash kernels."""
    from textboost_b200 import ops
    g = torch.Generator(device=dev).manual_seed(N + d)
    Cc = H * d
    qkv = torch.randn(B, N, 3 * Cc, device=dev, dtype=F16, generator=g)
    q, k, v = qkv[..., :Cc], qkv[..., Cc:2 * Cc], qkv[..., 2 * Cc:]
    do = torch.randn(B, N, Cc, device=dev, dtype=F16, generator=g)

    def heads(t):
        return t.reshape(B, -1, H, d).transpose(1, 2).float().detach().requires_grad_(True)

    qr, kr, vr = heads(q), heads(k), heads(v)
    oref = F.scaled_dot_product_attention(qr, kr, vr, is_causal=True)
    oref.backward(do.reshape(B, N, H, d).transpose(1, 2).float())
    o, lse = ops.attn_fwd(q, k, v, H, causal=True)
    assert relerr(o, oref.transpose(1, 2).reshape(B, N, Cc)) < TOL_ATTN
    dq, dk, dv = ops.attn_bwd(q, k, v, o, do, lse, H, causal=True)
    assert relerr(dq, qr.grad.transpose(1, 2).reshape(B, N, Cc)) < TOL_ATTN
    assert relerr(dk, kr.grad.transpose(1, 2).reshape(B, N, Cc)) < TOL_ATTN
    assert relerr(dv, vr.grad.transpose(1, 2).reshape(B, N, Cc)) < TOL_ATTN
    if N <= 128:  # the text encoder's case: dQ written as fp16 straight into the fused [dq | dk | dv] tensor
        dqkv = torch.zeros_like(qkv)
        dq16, _, _ = ops.attn_bwd(q, k, v, o, do, lse, H, causal=True, dk=dqkv[..., Cc:2 * Cc], dv=dqkv[..., 2 * Cc:],
                                  dq_out=dqkv[..., :Cc])
        assert relerr(dqkv[..., :Cc], qr.grad.transpose(1, 2).reshape(B, N, Cc)) < TOL_ATTN
        assert relerr(dqkv[..., Cc:2 * Cc], dk) < 1e-3 and relerr(dqkv[..., 2 * Cc:], dv) < 1e-3


@pytest.mark.parametrize("M,D,R,RPAD", [(616, 768, 12, 16), (77, 1024, 48, 48), (19, 768, 4, 16), (1232, 768, 64, 64),
                                        (40, 128, 12, 16)])
def test_text_encoder_layernorm_with_lora_glue(M, D, R, RPAD):
    """tb_layernorm_lora_fwd / tb_layernorm_bwd_clip against torch: LayerNorm on the fp32 residual stream with the LoRA
    down-projection (forward) and its input-gradient + the fp16 copy of dx (backward) in the same launch."""
    from textboost_b200 import ops
    g = torch.Generator(device=dev).manual_seed(M + D + R)
    x = torch.randn(M, D, device=dev, generator=g) * 2 + 0.3
    gamma = torch.randn(D, device=dev, generator=g)
    beta = torch.randn(D, device=dev, generator=g)
    A = torch.randn(R, D, device=dev, generator=g) / R
    y_ext = torch.full((M, D + RPAD + 8), 5.0, device=dev, dtype=F16)
    stats = ops.layernorm_lora_fwd(x, gamma, beta, A, y_ext, RPAD)
    yref = F.layer_norm(x, (D,), gamma, beta)
    assert relerr(y_ext[:, :D], yref) < 1e-3
    xa = y_ext[:, :D].float() @ A.t()
    assert relerr(y_ext[:, D:D + R], xa) < 2e-3
    assert bool((y_ext[:, D + R:D + RPAD] == 0).all()) and bool((y_ext[:, D + RPAD:] == 5.0).all())
    torch.testing.assert_close(stats[:, 0], x.mean(-1), rtol=1e-4, atol=1e-5)
    # backward: dy_ext = [dy | dxa] fp16, residual-stream gradient g32 added, fp16 copy written
    dy_ext = torch.randn(M, D + RPAD, device=dev, dtype=F16, generator=g)
    add = torch.randn(M, D, device=dev, generator=g)
    xr = x.clone().requires_grad_(True)
    dy_eff = dy_ext[:, :D].float() + dy_ext[:, D:D + R].float() @ A
    F.layer_norm(xr, (D,), gamma, beta).backward(dy_eff)
    out16 = torch.empty(M, D, device=dev, dtype=F16)
    buf = add.clone()
    dx = ops.layernorm_bwd_clip(dy_ext, x, gamma, stats, add=buf, out=buf, out16=out16, lora_a=A)
    assert dx.data_ptr() == buf.data_ptr()
    torch.testing.assert_close(dx, xr.grad + add, rtol=2e-4, atol=2e-4)
    torch.testing.assert_close(out16.float(), dx, rtol=1e-3 * TOLX, atol=1e-3 * TOLX)
    # plain variants: fp32 dy (final LayerNorm), no LoRA, no add, no fp16 copy
    dy32 = torch.randn(M, D, device=dev, generator=g)
    xr2 = x.clone().requires_grad_(True)
    F.layer_norm(xr2, (D,), gamma, beta).backward(dy32)
    torch.testing.assert_close(ops.layernorm_bwd_clip(dy32, x, gamma, stats), xr2.grad, rtol=2e-4, atol=2e-4)
    dx3 = ops.layernorm_bwd_clip(dy_ext[:, :D], x, gamma, stats, out16=out16)
    xr3 = x.clone().requires_grad_(True)
    F.layer_norm(xr3, (D,), gamma, beta).backward(dy_ext[:, :D].float())
    torch.testing.assert_close(dx3, xr3.grad, rtol=2e-4, atol=2e-4)


@pytest.mark.parametrize("B,HW,Cc,silu", [(2, 64, 64, True), (2, 4096, 320, True), (3, 1024, 640, False),
                                          (2, 256, 1920, True), (2, 64, 2560, True),
                                          # the single-launch group-owner kernel at the UNet's batch-8 shapes
                                          (8, 1024, 640, True), (8, 256, 1280, True), (8, 64, 1280, False),
                                          (8, 1024, 1280, True), (8, 256, 1920, True), (8, 64, 2560, True),
                                          # the cluster variant of it (slab split over 2 / 4 / 8 CTAs, totals through
                                          # distributed shared memory): the 64x64 level, the wide 32x32 concatenations,
                                          # the 96x96 level of a 768^2 image
                                          (8, 4096, 320, True), (8, 4096, 640, False), (8, 4096, 960, True),
                                          (8, 1024, 1920, True), (8, 9216, 320, False)])
def test_groupnorm_fwd_bwd(B, HW, Cc, silu):
    from textboost_b200 import ops
    torch.manual_seed(HW + Cc)
    x = (torch.randn(B, HW, Cc, device=dev) * 1.5 + 0.3).to(F16)
    gamma = (1 + 0.1 * torch.randn(Cc, device=dev)).to(F16)
    beta = (0.1 * torch.randn(Cc, device=dev)).to(F16)
    dy = torch.randn(B, HW, Cc, device=dev, dtype=F16)
    add = torch.randn(B, HW, Cc, device=dev, dtype=F16)
    xr = x.float().transpose(1, 2).requires_grad_(True)
    yr = F.group_norm(xr, 32, gamma.float(), beta.float(), 1e-5)
    if silu:
        yr = F.silu(yr)
    yr.backward(dy.float().transpose(1, 2))
    y, st = ops.groupnorm(x, gamma, beta, 32, 1e-5, silu)
    assert relerr(y, yr.transpose(1, 2)) < 2e-3
    dx = ops.groupnorm_bwd(dy, x, gamma, beta, st, 32, 1e-5, silu, add=add)
    assert relerr(dx, xr.grad.transpose(1, 2) + add.float()) < 3e-3
    # the saved statistics are (sum x, sum x^2) per (image, group) whichever kernel wrote them
    xs = x.float().view(B, HW, 32, Cc // 32)
    torch.testing.assert_close(st[..., 0], xs.sum((1, 3)), rtol=1e-4, atol=1e-2 * HW ** 0.5)
    torch.testing.assert_close(st[..., 1], (xs * xs).sum((1, 3)), rtol=1e-4, atol=1e-2 * HW ** 0.5)


def test_groupnorm_cluster_backward_and_8_cta_variants_in_a_child_process():
    """The cluster kernel also has a backward mode and 8-CTA clusters, off by default because the two-pass kernels
    measure faster there (csrc/norm.cu gn_cluster_geom); TB_GN_CLUSTER_ALL=1 (read once per process) turns them on:
    the same GroupNorm cases must pass through them."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, TB_GN_CLUSTER_ALL="1", TB_GN_CLUSTER_THREADS="512", TB_GN_CLUSTER_SMEM_KB="200")
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_kernels.py", "-m", "gpu", "-q", "-k",
                        "test_groupnorm_fwd_bwd", "-p", "no:cacheprovider"], cwd=root, env=env, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0 and " passed" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize("M,Cc,f32", [(512, 320, False), (77, 1280, False), (616, 768, True), (154, 1024, True),
                                      (130, 320, False), (1001, 640, False), (3, 640, False), (37, 128, True)])
def test_layernorm_fwd_bwd(M, Cc, f32):
    from textboost_b200 import ops
    torch.manual_seed(M + Cc)
    dt = torch.float32 if f32 else F16
    x = (torch.randn(M, Cc, device=dev) * 2 + 0.5).to(dt)
    gamma = (1 + 0.1 * torch.randn(Cc, device=dev)).to(dt)
    beta = (0.1 * torch.randn(Cc, device=dev)).to(dt)
    dy = torch.randn(M, Cc, device=dev, dtype=F16)
    add = torch.randn(M, Cc, device=dev, dtype=dt)
    xr = x.float().requires_grad_(True)
    yr = F.layer_norm(xr, (Cc,), gamma.float(), beta.float(), 1e-5)
    yr.backward(dy.float())
    y, st = ops.layernorm(x, gamma, beta)
    assert relerr(y, yr) < 2e-3
    dx = ops.layernorm_bwd(dy, x, gamma, st, add=add)
    assert relerr(dx, xr.grad + add.float()) < (1e-5 if f32 else 3e-3)


def test_elementwise_ops():
    from textboost_b200 import ops
    torch.manual_seed(3)
    h = torch.randn(300, 2 * 640, device=dev, dtype=F16)
    hr = h.float().requires_grad_(True)
    a, gate = hr.chunk(2, -1)
    gr = a * F.gelu(gate)
    dg = torch.randn(300, 640, device=dev, dtype=F16)
    gr.backward(dg.float())
    assert relerr(ops.geglu(h), gr) < 2e-3
    assert relerr(ops.geglu_bwd(dg, h), hr.grad) < 3e-3
    t = torch.tensor([0, 1, 500, 999], device=dev)
    te = ops.timestep_embedding(t, 320)
    fr = torch.exp(-math.log(10000.0) * torch.arange(160, device=dev).float() / 160)
    ref = torch.cat([torch.cos(t[:, None].float() * fr), torch.sin(t[:, None].float() * fr)], -1)
    assert (te.float() - ref).abs().max() < 2e-3
    # DDPM add_noise / velocity target (train_textboost.py:1052, 1070-1075)
    from textboost_b200.trainer import alphas_cumprod
    acp = alphas_cumprod(device=dev)
    x0, eps = torch.randn(4, 4, 16, 16, device=dev), torch.randn(4, 4, 16, 16, device=dev)
    sa, sb = acp[t].sqrt().view(-1, 1, 1, 1), (1 - acp[t]).sqrt().view(-1, 1, 1, 1)
    noisy, target = ops.add_noise(x0, eps, t, acp, v_prediction=True)
    assert relerr(noisy, sa * x0 + sb * eps) < 1e-3
    torch.testing.assert_close(target, sa * eps - sb * x0, rtol=1e-5, atol=1e-6)
    noisy, target = ops.add_noise(x0, eps, t, acp, v_prediction=False)
    assert torch.equal(target, eps)
    # MSE forward + gradient with a loss scale
    pred = torch.randn(4, 4, 16, 16, device=dev, dtype=F16)
    loss = torch.zeros(1, device=dev)
    scale = torch.full((1,), 128.0, device=dev)
    dpred = ops.mse_fwd_bwd(pred, target, loss, 1.0, scale)
    torch.testing.assert_close(loss[0], F.mse_loss(pred.float(), target), rtol=1e-5, atol=1e-6)
    assert relerr(dpred, 128.0 * 2 * (pred.float() - target) / pred.numel()) < 1e-3


def test_conv_in_out():
    from textboost_b200 import ops
    torch.manual_seed(4)
    x = torch.randn(2, 4, 16, 16, device=dev, dtype=F16)
    w = torch.randn(64, 4, 3, 3, device=dev, dtype=F16) / 6
    b = torch.randn(64, device=dev, dtype=F16)
    y = ops.conv_in(x, w, b)
    assert relerr(y, F.conv2d(x.float(), w.float(), b.float(), padding=1).permute(0, 2, 3, 1)) < TOL_GEMM
    h = torch.randn(2, 16, 16, 64, device=dev, dtype=F16)
    wo = torch.randn(4, 64, 3, 3, device=dev, dtype=F16) / 24
    bo = torch.randn(4, device=dev, dtype=F16)
    hr = h.float().permute(0, 3, 1, 2).requires_grad_(True)
    ref = F.conv2d(hr, wo.float(), bo.float(), padding=1)
    assert relerr(ops.conv_out(h, wo, bo), ref) < TOL_GEMM
    dy = torch.randn(2, 4, 16, 16, device=dev, dtype=F16)
    ref.backward(dy.float())
    assert relerr(ops.conv_out_bwd(dy, wo), hr.grad.permute(0, 2, 3, 1)) < TOL_GEMM


def test_fused_adamw_matches_torch():
    """tb_adamw_fused_step == GradScaler.unscale_ + clip_grad_norm_(LoRA only) + torch.optim.AdamW + renorm
    (train_textboost.py:1128-1149), over several steps, including an inf step that must be skipped."""
    from textboost_b200 import _cabi as C
    torch.manual_seed(5)
    n_lora, n_rows, D = 4096, 3, 64
    n = n_lora + n_rows * D
    p = torch.randn(n, device=dev) * 0.05
    lora_ref = torch.nn.Parameter(p[:n_lora].clone())
    rows_ref = torch.nn.Parameter(p[n_lora:].clone().view(n_rows, D))
    opt = torch.optim.AdamW([{"params": [rows_ref], "lr": 1e-3}, {"params": [lora_ref]}], lr=1e-4,
                            betas=(0.9, 0.999), weight_decay=1e-2, eps=1e-8)
    m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    state = torch.zeros(16, device=dev)
    state[0], state[5] = 1024.0, 1.0
    mean_norm = 0.3
    norm_out = torch.zeros(1, device=dev)
    world = 2
    for step in range(6):
        g = torch.randn(n, device=dev) * (3.0 if step == 1 else 0.01)
        scale = state[0].item()
        grads = g * scale * world  # as if summed over `world` ranks holding the same gradient
        if step == 3:
            grads[7] = float("inf")
        before = p.clone()
        C.call("tb_adamw_fused_step", C.ptr(p), C.ptr(grads), C.ptr(m), C.ptr(v), n_lora, n_rows, D, 1e-4, 1e-3,
               0.9, 0.999, 1e-8, 1e-2, 1.0, 1.0 / world, mean_norm, 0, 0.0, 0.0, C.ptr(state), C.ptr(norm_out),
               C.stream_ptr())
        torch.cuda.synchronize()
        assert torch.count_nonzero(grads) == 0  # zero_grad fused
        if step == 3:
            assert torch.equal(p[:n_lora], before[:n_lora]) and state[0].item() == scale / 2 and state[8].item() == 1
            continue
        lora_ref.grad, rows_ref.grad = g[:n_lora].clone(), g[n_lora:].clone().view(n_rows, D)
        gn = torch.nn.utils.clip_grad_norm_([lora_ref], 1.0)
        opt.step()
        with torch.no_grad():
            nv = rows_ref.norm(dim=-1, keepdim=True)
            rows_ref.copy_(torch.minimum(torch.full_like(nv, mean_norm), nv) / nv * rows_ref)
        assert abs(state[7].item() - gn.item()) < 1e-4 * gn.item()
        torch.testing.assert_close(p[:n_lora], lora_ref.detach(), rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(p[n_lora:].view(n_rows, D), rows_ref.detach(), rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(norm_out[0], nv.mean(), rtol=1e-5, atol=1e-7)
    assert state[4].item() == 5 and abs(state[5].item() - (1 - 1e-3 * 1e-2) ** 5) < 1e-6


@pytest.mark.parametrize("name", ["constant_with_warmup", "linear", "cosine", "cosine_with_restarts", "polynomial"])
def test_fused_adamw_lr_schedules_match_torch_lambda_lr(name):
    """--lr_scheduler (train_textboost.py:224-233, 911-916): the schedule evaluated on the device inside
    tb_adamw_fused_step against torch.optim.AdamW driven by a LambdaLR with the same diffusers formula, 12 steps with
    3 warm-up steps; a GradScaler-skipped step (inf gradient) does not advance the schedule."""
    from textboost_b200 import _cabi as C
    from textboost_b200.optim import LR_SCHEDULES, lr_multiplier
    n_lora, D, n_rows, warm, total = 64, 16, 1, 3, 12
    n = n_lora + n_rows * D
    g0 = torch.Generator(device=dev).manual_seed(5)
    p = torch.randn(n, device=dev, generator=g0)
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([ref], lr=1e-2, weight_decay=1e-2)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: lr_multiplier(name, s, warm, total, 1e-2))
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    state = torch.zeros(16, device=dev)
    state[0], state[5] = 1.0, 1.0
    norm_out = torch.zeros(1, device=dev)
    for step in range(total):
        g = torch.randn(n, device=dev, generator=g0) * 0.1
        grads = g.clone()
        if step == 5:
            grads[3] = float("inf")
        C.call("tb_adamw_fused_step", C.ptr(p), C.ptr(grads), C.ptr(m), C.ptr(v), n_lora, n_rows, D, 1e-2, 1e-2,
               0.9, 0.999, 1e-8, 1e-2, 0.0, 1.0, 1e9, LR_SCHEDULES[name], float(warm), float(total), C.ptr(state),
               C.ptr(norm_out), C.stream_ptr())
        if step == 5:
            state[0] = 1.0  # undo the GradScaler back-off so the comparison stays at scale 1
            continue
        ref.grad = g.clone()
        opt.step()
        sched.step()
        assert abs(state[9].item() - lr_multiplier(name, int(state[4].item()) - 1, warm, total, 1e-2)) < 1e-6
    torch.testing.assert_close(p, ref.detach(), rtol=2e-5, atol=2e-6)
    assert state[4].item() == total - 1 and state[8].item() == 1


@pytest.mark.parametrize("r", [4, 8])
def test_unet_cross_kv_lora_fwd_bwd(r):
    """tb_unet_lora_fwd / tb_unet_lora_bwd (peft LoRA on attn2.to_k / to_v of all blocks at once,
    train_textboost.py:712-721) against autograd on the same fp32 adapters: three blocks of different widths."""
    from textboost_b200 import ops
    g = torch.Generator(device=dev).manual_seed(5 + r)
    M, ctx, widths = 154, 128, (64, 128, 64)
    KV = 2 * sum(widths)
    n_ad = 2 * len(widths)
    R = n_ad * r
    blk, off = [], [0]
    for i, w in enumerate(widths):
        blk += [2 * i] * w + [2 * i + 1] * w
        off += [off[-1] + w, off[-1] + 2 * w]
    blk_t = torch.tensor(blk, dtype=torch.int32, device=dev)
    off_t = torch.tensor(off, dtype=torch.int32, device=dev)
    ehs = torch.randn(M, ctx, device=dev, generator=g).to(F16)
    kv0 = torch.randn(M, KV, device=dev, generator=g).to(F16)
    A = (torch.randn(R, ctx, device=dev, generator=g) / r).requires_grad_(True)
    Bm = (0.05 * torch.randn(KV, r, device=dev, generator=g)).requires_grad_(True)
    dkv = torch.randn(M, KV, device=dev, generator=g).to(F16)
    scaling = 1.5
    # reference: per adapter, kv[:, cols] += scaling * (ehs A_a^T) B_a^T
    e32 = ehs.float().requires_grad_(True)
    parts = []
    for a in range(n_ad):
        lo, hi = off[a], off[a + 1]
        parts.append(scaling * (e32 @ A[a * r:(a + 1) * r].t()) @ Bm[lo:hi].t())
    ref = kv0.float() + torch.cat(parts, 1)
    ref.backward(dkv.float())
    kv = kv0.clone()
    Z = ops.unet_lora_fwd(ehs, A.detach(), Bm.detach(), blk_t, kv, r, scaling)
    assert relerr(kv, ref) < 2e-3
    torch.testing.assert_close(Z, ehs.float() @ A.detach().t(), rtol=1e-4, atol=1e-4)
    dA = torch.full_like(A, 0.25)  # accumulating outputs
    dB = torch.full_like(Bm, -0.5)
    d_ehs = torch.ones(M, ctx, device=dev)
    ops.unet_lora_bwd(dkv, ehs, A.detach(), Bm.detach(), Z, blk_t, off_t, dA, dB, d_ehs, r, scaling)
    torch.testing.assert_close(dA - 0.25, A.grad, rtol=2e-4, atol=2e-4 * A.grad.abs().max().item())
    torch.testing.assert_close(dB + 0.5, Bm.grad, rtol=2e-4, atol=2e-4 * Bm.grad.abs().max().item())
    torch.testing.assert_close(d_ehs - 1.0, e32.grad, rtol=2e-4, atol=2e-4 * e32.grad.abs().max().item())
