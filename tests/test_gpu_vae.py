"""GPU parity of the AutoencoderKL encoder engine (image half) against oracle/vae_ref.py, through
the C ABI.  The oracle runs in fp32 (the reference keeps the VAE in fp32, train_textboost.py:938); the engine computes
in fp16 with fp32 accumulation, so the tolerances below are fp16 ones: relative L2 <= 1e-2 on the posterior mean /
std and on the scaled latents.  The helper kernels are checked exactly (window gather) or to fp16 rounding."""
import pytest
import torch
import torch.nn.functional as F

from conftest import act_dtype, tol_scale

pytestmark = pytest.mark.gpu
F16 = act_dtype()  # fp16; bf16 when the file is re-run under the bf16 policy (tests/test_gpu_bf16_policy.py)
TOLX = tol_scale()  # 1 for fp16, 8 for bf16: rel_l2() reports errors in units of the fp16 bounds written below
dev = "cuda"


@pytest.fixture(scope="module", autouse=True)
def _need_gpu(built_lib):
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from textboost_b200 import _cabi
    _cabi.call("tb_check_device")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def rel_l2(a, b):
    return (((a.float() - b.float()).norm() / (b.float().norm() + 1e-20)).item()) / TOLX


@pytest.mark.parametrize("pad_lo", [0, 1])
@pytest.mark.parametrize("shape", [(2, 16, 24, 64), (1, 6, 4, 8)])
def test_im2col_stride2_padding_variants_exact(shape, pad_lo):
    from textboost_b200 import ops
    B, H, W, Cc = shape
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, H, W, Cc, generator=g).to(F16).to(dev)
    col = ops.im2col3x3s2_pad(x, pad_lo)
    xn = x.permute(0, 3, 1, 2).float()
    xp = F.pad(xn, (pad_lo, 1 - pad_lo + 1, pad_lo, 1 - pad_lo + 1))  # enough zeros on the high side for both
    ref = torch.empty(B, H // 2, W // 2, 9, Cc, device=dev)
    for ky in range(3):
        for kx in range(3):
            ref[:, :, :, ky * 3 + kx] = xp[:, :, ky:ky + H:2, kx:kx + W:2].permute(0, 2, 3, 1)
    assert torch.equal(col.float(), ref.reshape(B * (H // 2) * (W // 2), 9 * Cc))
    if pad_lo == 1:
        assert torch.equal(col, ops.im2col3x3s2(x))


def test_vae_downsample_matches_padded_strided_conv():
    """window gather (pad_lo=0) + GEMM == F.conv2d(F.pad(x, (0,1,0,1)), w, stride=2) (diffusers Downsample2D)."""
    from textboost_b200 import ops
    from textboost_b200.unet import _conv_fwd_weight
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 64, 16, 16, generator=g).to(F16).to(dev)
    w = (torch.randn(128, 64, 3, 3, generator=g) * 0.05).to(F16).to(dev)
    b = torch.randn(128, generator=g).to(F16).to(dev)
    ref = F.conv2d(F.pad(x.float(), (0, 1, 0, 1)), w.float(), b.float(), stride=2)
    col = ops.im2col3x3s2_pad(x.permute(0, 2, 3, 1).contiguous(), 0)
    y = ops.gemm(col, _conv_fwd_weight(w), bias=b).view(2, 8, 8, 128).permute(0, 3, 1, 2)
    assert rel_l2(y, ref) < 2e-3


@pytest.mark.parametrize("rows,cols,ld", [(64, 64, 64), (300, 1024, 1024), (128, 4096, 4096), (5, 8192, 8192),
                                          (33, 72, 80)])
def test_softmax_rows(rows, cols, ld):
    from textboost_b200 import ops
    g = torch.Generator().manual_seed(rows)
    buf = (torch.randn(rows, ld, generator=g) * 4).to(F16).to(dev)
    x = buf[:, :cols]
    ref = torch.softmax(x.float(), dim=-1)
    keep = buf[:, cols:].clone()
    ops.softmax_rows_(x)
    assert (x.float() - ref).abs().max().item() < 1e-3 * TOLX and rel_l2(x, ref) < 1e-3
    assert (x.float().sum(-1) - 1).abs().max().item() < 5e-3 * TOLX
    assert torch.equal(buf[:, cols:], keep)  # the padding columns of a strided view stay untouched


def test_vae_sample_kernel():
    from textboost_b200 import ops
    g = torch.Generator().manual_seed(5)
    B, HW, L = 3, 48, 4
    rows = torch.randn(B * HW, 64, generator=g)
    rows[:, L:2 * L] *= 20  # exercise the logvar clamp at +20 (the -30 side underflows to std ~ 3e-7)
    rows = rows.to(F16).to(dev)
    eps = torch.randn(B, L, HW, generator=g).to(dev)
    lat, mean, std = ops.vae_sample(rows, B, HW, L, eps=eps, scaling_factor=0.18215, want_moments=True)
    m = rows[:, :L].float().view(B, HW, L).transpose(1, 2)
    s = torch.exp(0.5 * rows[:, L:2 * L].float().clamp(-30, 20)).view(B, HW, L).transpose(1, 2)
    assert torch.equal(mean, m.contiguous())
    assert rel_l2(std, s) < 1e-6
    assert rel_l2(lat, (m + s * eps) * 0.18215) < 1e-6


def _oracle_and_engine(seed, cfg_kwargs=None):
    from oracle import vae_ref
    from textboost_b200 import vae
    torch.manual_seed(seed)
    ocfg = vae_ref.VAEConfig(**(cfg_kwargs or {}))
    ref = vae_ref.AutoencoderKLEncoderRef(ocfg)
    with torch.no_grad():  # non-trivial norm affine parameters and biases
        for n, p in ref.named_parameters():
            if "norm" in n and n.endswith("weight"):
                p.add_(0.1 * torch.randn_like(p))
            elif n.endswith("bias"):
                p.add_(0.05 * torch.randn_like(p))
    ref = ref.to(dev).eval().requires_grad_(False)
    eng = vae.VAEEncoderEngine(vae.VAEConfig(**(cfg_kwargs or {})), ref.state_dict())
    return ref, eng


@pytest.mark.parametrize("B,H,W", [(2, 256, 256), (1, 512, 512), (5, 64, 96)])
def test_vae_encoder_engine_matches_oracle_full_config(B, H, W):
    """SD-1.x / 2.x VAE encoder (34.2 M parameters) at the training resolution and two smaller ones (ragged chunk:
    5 images over max_chunk 4; non-square 64x96)."""
    ref, eng = _oracle_and_engine(11)
    g = torch.Generator().manual_seed(B * H)
    px = (torch.rand(B, 3, H, W, generator=g) * 2 - 1).to(dev)
    eps = torch.randn(B, 4, H // 8, W // 8, generator=g).to(dev)
    with torch.no_grad():
        mean_r, std_r = ref.moments(px)
        lat_r = ref.encode_latents(px, eps)
    mean, std = eng.moments(px)
    lat = eng.encode_latents(px, eps)
    assert mean.shape == mean_r.shape and lat.shape == lat_r.shape and lat.dtype == torch.float32
    assert torch.isfinite(lat).all()
    assert rel_l2(mean, mean_r) < 1e-2, rel_l2(mean, mean_r)
    assert rel_l2(std, std_r) < 1e-2, rel_l2(std, std_r)
    assert rel_l2(lat, lat_r) < 1e-2, rel_l2(lat, lat_r)
    # run to run only the order of GroupNorm's fp32 atomic partial sums differs
    assert rel_l2(lat, eng.encode_latents(px, eps)) < 2e-3
    print(f"VAE_PARITY B={B} {H}x{W} mean {rel_l2(mean, mean_r):.2e} std {rel_l2(std, std_r):.2e} "
          f"latents {rel_l2(lat, lat_r):.2e}")


def test_autoencoder_kl_mirror_reference_call_sequence(tmp_path):
    """vae = AutoencoderKL.from_pretrained(path, subfolder="vae"); vae.eval().requires_grad_(False);
    vae.to(device, dtype=torch.float32); vae.encode(px.to(dtype=vae.dtype)).latent_dist.sample() * scaling_factor
    (train_textboost.py:651-653, 697, 938, 1027, 1036-1037) on a small checkpoint in the diffusers layout."""
    import json
    import os
    from safetensors.torch import save_file
    from textboost_b200.vae import AutoencoderKL
    kw = dict(block_out_channels=(64, 64, 128, 128))
    ref, _ = _oracle_and_engine(13, kw)
    d = tmp_path / "ckpt" / "vae"
    os.makedirs(d)
    save_file({k: v.cpu().contiguous() for k, v in ref.state_dict().items()},
              str(d / "diffusion_pytorch_model.safetensors"))
    with open(d / "config.json", "w") as f:
        json.dump({"_class_name": "AutoencoderKL", "in_channels": 3, "latent_channels": 4,
                   "block_out_channels": [64, 64, 128, 128], "layers_per_block": 2, "norm_num_groups": 32,
                   "scaling_factor": 0.18215}, f)
    vae = AutoencoderKL.from_pretrained(str(tmp_path / "ckpt"), subfolder="vae", revision=None, variant=None)
    vae.eval().requires_grad_(False)
    with pytest.raises(RuntimeError):
        vae.encode(torch.zeros(1, 3, 64, 64))
    vae.to(dev, dtype=torch.float32)
    assert vae.config.scaling_factor == 0.18215 and vae.dtype == torch.float32
    px = (torch.rand(2, 3, 128, 128, generator=torch.Generator().manual_seed(1)) * 2 - 1).to(dev, dtype=vae.dtype)
    dist = vae.encode(px).latent_dist
    with torch.no_grad():
        mean_r, std_r = ref.moments(px)
    assert rel_l2(dist.mode(), mean_r) < 1e-2 and rel_l2(dist.std, std_r) < 1e-2
    gen = torch.Generator(device=dev).manual_seed(7)
    lat = dist.sample(generator=gen) * vae.config.scaling_factor
    gen.manual_seed(7)
    eps = torch.randn(mean_r.shape, device=dev, generator=gen)
    assert rel_l2(lat, (mean_r + std_r * eps) * 0.18215) < 1e-2
    with pytest.raises(ValueError):
        vae.encode(torch.zeros(1, 3, 60, 64, device=dev)).latent_dist.mode()
