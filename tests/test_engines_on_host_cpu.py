"""The forward-only engines on the CPU through the product's SIMT kernels (host build of the C ABI) and the torch
contracts of the tensor-core wrappers (tests/engine_standin.py): AutoencoderKL encoder / decoder engines and the sampling
loop with the real UNetEngine, against the oracles — the CPU-side counterpart of tests/test_gpu_vae.py and
tests/test_gpu_sampler.py."""
from types import SimpleNamespace

import pytest
import torch

import engine_standin

pytestmark = pytest.mark.timeout(900, method="thread")
KW = dict(block_out_channels=(64, 64, 128, 128))


def rel_l2(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-20)).item()


def test_vae_engines_through_the_real_simt_kernels(monkeypatch):
    from oracle import vae_ref
    from textboost_b200 import vae
    engine_standin.install(monkeypatch)
    monkeypatch.setattr(vae.VAEEncoderEngine, "_check", lambda self, p: None)
    monkeypatch.setattr(vae.VAEDecoderEngine, "_check", lambda self, p: None)
    torch.manual_seed(0)
    ref = vae_ref.AutoencoderKLRef(vae_ref.VAEConfig(**KW)).eval().requires_grad_(False)
    enc = vae.VAEEncoderEngine(vae.VAEConfig(**KW), ref.state_dict())
    dec = vae.VAEDecoderEngine(vae.VAEConfig(**KW), ref.state_dict())
    px = torch.rand(2, 3, 32, 32) * 2 - 1
    eps = torch.randn(2, 4, 4, 4)
    mean_r, std_r = ref.moments(px)
    mean, std = enc.moments(px)
    assert rel_l2(mean, mean_r) < 5e-3 and rel_l2(std, std_r) < 5e-3
    assert rel_l2(enc.encode_latents(px, eps), ref.encode_latents(px, eps)) < 5e-3
    lat = torch.randn(2, 4, 4, 4) * 0.18215 * 3
    assert rel_l2(dec.decode(lat, scaling_factor=0.18215), ref.decode(lat / 0.18215)) < 5e-3
    d = (dec.decode_u8(lat).int() - ref.to_uint8(ref.decode_latents(lat)).int()).abs()
    assert d.max() <= 2


@pytest.mark.long_cpu
def test_sampling_loop_with_the_real_unet_engine(monkeypatch):
    """StableDiffusionPipeline.denoise (UNetEngine forward on the doubled batch + tb_dpm_cfg_step per step) vs the
    oracle sampler driving the oracle UNet built from the same weights: 5 steps, guidance 7.5."""
    from oracle import harness, sampler_ref, unet_ref
    from textboost_b200 import pipeline, synthetic
    from textboost_b200.unet import UNetEngine
    engine_standin.install(monkeypatch)
    ucfg, _ = synthetic.model_configs("tiny")
    usd = synthetic.random_unet_sd(ucfg, "cpu", 3)
    eng = UNetEngine(ucfg, usd)
    oracle = unet_ref.UNet2DConditionModelRef(harness._unet_cfg(ucfg))
    oracle.load_state_dict({k: v.float() for k, v in usd.items()})
    oracle.eval().requires_grad_(False)
    unet = SimpleNamespace(engine=eng, config=SimpleNamespace(sample_size=8, in_channels=4))
    pipe = pipeline.StableDiffusionPipeline(SimpleNamespace(config={"block_out_channels": (1, 2, 3, 4)}), None, None, unet,
                                            pipeline.DPMSolverMultistepScheduler())
    pipe.use_cuda_graph = False
    g = torch.Generator().manual_seed(4)
    cond = (torch.randn(2, 77, ucfg.cross_attention_dim, generator=g) * 0.5).half()
    uncond = (torch.randn(2, 77, ucfg.cross_attention_dim, generator=g) * 0.5).half()
    lat = pipe.prepare_latents(2, 64, 64, "cpu", generator=torch.Generator().manual_seed(5))
    x = pipe.denoise(lat.clone(), cond, uncond, 5, 7.5)
    ref = sampler_ref.sample_latents(lambda a, t, e: oracle(a, t, e), cond.float(), uncond.float(), lat.clone(),
                                     sampler_ref.DPMSolverMultistepRef(), 5, 7.5)
    assert torch.isfinite(x).all() and rel_l2(x, ref) < 2e-2, rel_l2(x, ref)
