"""TEST INFRASTRUCTURE ONLY: plain-PyTorch CPU stand-ins for the `textboost_b200.ops` wrappers that the forward-only
engines (VAE encoder / decoder, sampler) call, so their ORCHESTRATION (layer order, weight layouts, folded weights,
chunking, views and strides) can be checked against the oracle in the CPU suite.  They keep the fp16 storage rounding of
the real kernels (fp32 math, fp16 outputs).  Nothing outside tests/ imports this; the product path has no CPU route —
`install(monkeypatch)` swaps the functions for the duration of one test."""
import torch
import torch.nn.functional as F

H16, F32 = torch.float16, torch.float32


def gemm(a, w, *, bias=None, rowvec=None, rows_per_group=1, residual=None, alpha=1.0, act=0, out=None, out_kind=0):
    assert a.dtype == H16 and w.dtype == H16 and a.stride(1) == 1 and w.stride(1) == 1 and act == 0
    y = a.float() @ w.float().t() * alpha
    if bias is not None:
        y = y + bias.float()
    if residual is not None:
        y = y + residual.float()
    y = y.half()
    if out is not None:
        out.copy_(y)
        return out
    return y


def conv3x3(x, w, *, bias=None, rowvec=None, residual=None, act=0, out=None):
    B, Hh, W, Cin = x.shape
    Cout = w.shape[0]
    assert x.dtype == H16 and x.is_contiguous() and w.shape[1] == 9 * Cin and Cin % 64 == 0 and act == 0
    wt = w.float().view(Cout, 3, 3, Cin).permute(0, 3, 1, 2)
    y = F.conv2d(x.float().permute(0, 3, 1, 2), wt, None if bias is None else bias.float(), padding=1)
    y = y.permute(0, 2, 3, 1)
    if residual is not None:
        y = y + residual.float()
    return y.half().contiguous()


def groupnorm(x, gamma, beta, groups, eps, silu):
    B, C = x.shape[0], x.shape[-1]
    y = F.group_norm(x.float().reshape(B, -1, C).transpose(1, 2), groups, gamma.float(), beta.float(), eps)
    if silu:
        y = F.silu(y)
    return y.transpose(1, 2).reshape(x.shape).half().contiguous(), None


def conv_in(x_nchw, w, b):
    assert x_nchw.dtype == H16 and x_nchw.shape[1] <= 8 and w.shape[0] % 8 == 0
    assert w.shape[0] * w.shape[1] * 9 * 2 <= 48 * 1024  # the kernel stages the whole weight table in shared memory
    return F.conv2d(x_nchw.float(), w.float(), b.float(), padding=1).permute(0, 2, 3, 1).half().contiguous()


def im2col3x3s2_pad(x, pad_lo):
    B, Hh, W, C = x.shape
    xp = F.pad(x.float().permute(0, 3, 1, 2), (pad_lo, 2 - pad_lo, pad_lo, 2 - pad_lo))
    col = torch.empty(B, Hh // 2, W // 2, 9, C)
    for ky in range(3):
        for kx in range(3):
            col[:, :, :, ky * 3 + kx] = xp[:, :, ky:ky + Hh:2, kx:kx + W:2].permute(0, 2, 3, 1)
    return col.reshape(B * (Hh // 2) * (W // 2), 9 * C).half()


def upsample2x(x):
    return x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2).contiguous()


def softmax_rows_(x):
    x.copy_(torch.softmax(x.float(), -1).half())
    return x


def vae_sample(m, B, HW, L, eps=None, scaling_factor=1.0, want_moments=False):
    mean = m[:, :L].float().view(B, HW, L).transpose(1, 2).contiguous()
    std = torch.exp(0.5 * m[:, L:2 * L].float().clamp(-30, 20)).view(B, HW, L).transpose(1, 2).contiguous()
    lat = (mean + std * eps.reshape(B, L, HW)) * scaling_factor if eps is not None else None
    return lat, (mean if want_moments else None), (std if want_moments else None)


def vae_decode_in(latents, w, bias, scaling_factor):
    z = torch.einsum("oc,bchw->bohw", w, latents / scaling_factor) + bias.view(1, -1, 1, 1)
    return z.half()


def image_u8(rows, npix, channels=3):
    v = (rows[:, :channels].float() * 0.5 + 0.5).clamp(0, 1)
    return (v * 255).round().to(torch.uint8)


def dpm_cfg_step(x, eps, m_prev, m_out, unet_in, guidance_scale, alpha_i, sigma_i, v_prediction, c_x, c_d0, c_d1):
    f = torch.float32
    g, a, s = (torch.tensor(v, dtype=f) for v in (guidance_scale, alpha_i, sigma_i))
    e_u, e_c = eps.float().view(2, -1)
    e = e_u + g * (e_c - e_u)
    xi = x.view(-1)
    m0 = a * xi - s * e if v_prediction else (xi - s * e) / a
    xn = torch.tensor(c_x, dtype=f) * xi + torch.tensor(c_d0, dtype=f) * m0
    if c_d1 != 0.0:
        xn = xn + torch.tensor(c_d1, dtype=f) * (m0 - m_prev.view(-1))
    m_out.view(-1).copy_(m0)
    xi.copy_(xn)
    if unet_in is not None:
        unet_in.view(2, -1).copy_(xn.half().expand(2, -1))
    return x


NAMES = ["gemm", "conv3x3", "groupnorm", "conv_in", "im2col3x3s2_pad", "upsample2x", "softmax_rows_", "vae_sample",
         "vae_decode_in", "image_u8", "dpm_cfg_step", "cast_f32_f16"]


def cast_f32_f16(src, out=None, scale=1.0):
    if out is None:
        out = torch.empty(src.shape, dtype=H16)
    out.copy_((src.float() * scale).to(H16))
    return out


def install(monkeypatch):
    from textboost_b200 import ops, vae
    for n in NAMES:
        monkeypatch.setattr(ops, n, globals()[n])
    # the engines refuse CPU tensors by design; lift that check for the orchestration tests only
    monkeypatch.setattr(vae.VAEEncoderEngine, "_check", lambda self, p: None)
    monkeypatch.setattr(vae.VAEDecoderEngine, "_check", lambda self, p: None)
