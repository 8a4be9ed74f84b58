"""Auto cross-check of the UNPINNED oracles against the real libraries — runs only where they import.

SURVEY.md §8c: diffusers / peft are the third-party modules that hold the reference's arithmetic for the UNet, the VAE,
the schedulers and the LoRA layers; neither is installed or installable in the build container, so every test here is
skipped there and the oracles that restate them say "parity unpinned".  On any box that has the pinned versions
(diffusers 0.29, peft 0.13) these tests load ONE state dict into both sides and compare forward results in fp32,
turning the restatements into pinned ones.  CPU only, small configurations, seconds each.
"""
import pytest
import torch

diffusers = pytest.importorskip("diffusers", reason="diffusers is not installed (expected in the build container)")


def _max_rel(a, b):
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def test_unet_oracle_matches_diffusers():
    from oracle import unet_ref
    cfg = unet_ref.UNetConfig.tiny(cross_attention_dim=64)
    torch.manual_seed(0)
    lib = diffusers.UNet2DConditionModel(
        sample_size=cfg.sample_size, in_channels=4, out_channels=4, block_out_channels=cfg.block_out_channels,
        layers_per_block=cfg.layers_per_block, cross_attention_dim=cfg.cross_attention_dim,
        attention_head_dim=cfg.attention_head_dim, norm_num_groups=cfg.norm_num_groups,
        down_block_types=("CrossAttnDownBlock2D",) * 3 + ("DownBlock2D",),
        up_block_types=("UpBlock2D",) + ("CrossAttnUpBlock2D",) * 3).eval()
    ours = unet_ref.UNet2DConditionModelRef(cfg).eval()
    ours.load_state_dict(lib.state_dict())
    x, t = torch.randn(2, 4, 16, 16), torch.tensor([10, 900])
    ehs = torch.randn(2, 77, 64, requires_grad=True)
    ehs2 = ehs.detach().clone().requires_grad_(True)
    a = lib(x, t, ehs).sample
    b = ours(x, t, ehs2)
    assert _max_rel(b, a) < 1e-5
    a.square().mean().backward()
    b.square().mean().backward()
    assert _max_rel(ehs2.grad, ehs.grad) < 1e-4


def test_vae_oracle_matches_diffusers():
    from oracle import vae_ref
    kw = dict(block_out_channels=(32, 32, 64, 64))
    torch.manual_seed(1)
    lib = diffusers.AutoencoderKL(in_channels=3, out_channels=3, latent_channels=4, layers_per_block=2,
                                  block_out_channels=kw["block_out_channels"], norm_num_groups=32,
                                  down_block_types=("DownEncoderBlock2D",) * 4,
                                  up_block_types=("UpDecoderBlock2D",) * 4).eval()
    ours = vae_ref.AutoencoderKLRef(vae_ref.VAEConfig(**kw)).eval()
    ours.load_state_dict(lib.state_dict())
    px = torch.rand(2, 3, 64, 64) * 2 - 1
    with torch.no_grad():
        dist = lib.encode(px).latent_dist
        mean, std = ours.moments(px)
        assert _max_rel(mean, dist.mean) < 1e-5 and _max_rel(std, dist.std) < 1e-5
        z = torch.randn(2, 4, 8, 8)
        assert _max_rel(ours.decode(z), lib.decode(z).sample) < 1e-5


@pytest.mark.parametrize("pred", ["epsilon", "v_prediction"])
@pytest.mark.parametrize("spacing,offset,n", [("linspace", 0, 25), ("leading", 1, 25), ("trailing", 0, 10)])
def test_dpm_solver_oracle_and_product_match_diffusers(pred, spacing, offset, n):
    from oracle import sampler_ref
    from textboost_b200.pipeline import DPMSolverMultistepScheduler
    kw = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
              prediction_type=pred, timestep_spacing=spacing, steps_offset=offset)
    lib = diffusers.DPMSolverMultistepScheduler(**kw)
    lib.set_timesteps(n)
    ref = sampler_ref.DPMSolverMultistepRef(prediction_type=pred, timestep_spacing=spacing, steps_offset=offset)
    ours = DPMSolverMultistepScheduler(**kw)
    assert ref.set_timesteps(n).tolist() == lib.timesteps.tolist() == ours.set_timesteps(n).tolist()
    g = torch.Generator().manual_seed(2)
    x = torch.randn(1, 4, 8, 8, generator=g, dtype=torch.float64) * float(lib.sigmas[0])
    xa, xb, xc, m_prev = x.clone(), x.clone(), x.clone(), None
    for i, t in enumerate(lib.timesteps):
        e = torch.randn(x.shape, generator=g, dtype=torch.float64)
        xa = lib.step(e, t, xa).prev_sample
        xb = ref.step(e, xb)
        k = ours.step_coefficients(i)
        m0 = k["alpha_i"] * xc - k["sigma_i"] * e if pred == "v_prediction" else (xc - k["sigma_i"] * e) / k["alpha_i"]
        xc = k["c_x"] * xc + k["c_d0"] * m0 + (k["c_d1"] * (m0 - m_prev) if k["c_d1"] else 0.0)
        m_prev = m0
        assert _max_rel(xb, xa) < 1e-5 and _max_rel(xc, xa) < 1e-5, i


def test_ddpm_oracle_matches_diffusers():
    from oracle import ddpm_ref
    lib = diffusers.DDPMScheduler(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012,
                                  beta_schedule="scaled_linear")
    torch.testing.assert_close(ddpm_ref.alphas_cumprod(), lib.alphas_cumprod, rtol=1e-6, atol=0)
    x0, eps, t = torch.randn(3, 4, 8, 8), torch.randn(3, 4, 8, 8), torch.tensor([0, 500, 999])
    torch.testing.assert_close(ddpm_ref.add_noise(x0, eps, t), lib.add_noise(x0, eps, t), rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(ddpm_ref.get_velocity(x0, eps, t), lib.get_velocity(x0, eps, t), rtol=1e-6, atol=1e-6)


def test_lora_oracle_matches_peft():
    peft = pytest.importorskip("peft", reason="peft is not installed")
    import copy
    from oracle import clip_ref
    cfg = clip_ref.ClipTextConfig(vocab_size=1000, hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                                  num_attention_heads=2)
    base = clip_ref.init_clip_(clip_ref.TextBoostModelRef(cfg), 0)
    ours = copy.deepcopy(base)
    ours.add_adapter(r=4)
    lib = peft.get_peft_model(copy.deepcopy(base), peft.LoraConfig(r=4, lora_alpha=4, init_lora_weights="gaussian",
                                                                   target_modules=["q_proj", "k_proj", "v_proj"]))
    sd = {n: p for n, p in ours.named_parameters() if "lora_" in n}
    with torch.no_grad():
        for n, p in lib.named_parameters():
            if "lora_" in n:
                key = n.replace("base_model.model.", "")
                p.copy_(sd[key] if "lora_A" in key else torch.randn_like(p) * 0.02)
                if "lora_B" in key:
                    sd[key].copy_(p)
    ids = torch.randint(0, 1000, (2, 77))
    assert _max_rel(ours(ids), lib(ids)) < 1e-5
