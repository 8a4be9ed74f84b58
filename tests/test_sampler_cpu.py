"""Validation / inference sampler, host side: the DPM-Solver++ schedule and per-step coefficients of
textboost_b200.pipeline against oracle/sampler_ref.py (diffusers DPMSolverMultistepScheduler restated; parity
unpinned), known values of the SD noise schedule, and the sampling loop's orchestration through the torch stand-ins of
tests/ops_standin.py."""
import json
import math
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import ops_standin


def test_sigma_schedule_known_values():
    """SD's scaled-linear schedule: sigma_max = 14.6146 at t = 999, sigma_min = 0.0292 at t = 0 (the k-diffusion
    constants), 25 linspace steps start 999, 959, 919; 'leading' with steps_offset 1 starts at 951 and ends at 39."""
    from oracle import sampler_ref
    from textboost_b200.pipeline import DPMSolverMultistepScheduler
    s = DPMSolverMultistepScheduler()
    ts = s.set_timesteps(25)
    assert ts[:3].tolist() == [999, 959, 919] and ts[-1] == 40 and len(ts) == 25 and len(s.sigmas) == 26
    assert abs(s.sigmas[0] - 14.6146) < 1e-3 and s.sigmas[-1] == 0.0
    acp = s.alphas_cumprod.double().numpy()
    assert abs(math.sqrt((1 - acp[0]) / acp[0]) - 0.0292) < 1e-4
    lead = DPMSolverMultistepScheduler(timestep_spacing="leading", steps_offset=1)
    assert lead.set_timesteps(25)[[0, -1]].tolist() == [951, 39]
    trail = DPMSolverMultistepScheduler(timestep_spacing="trailing")
    assert trail.set_timesteps(4).tolist() == [999, 749, 499, 249]
    for spacing, off, n in (("linspace", 0, 25), ("linspace", 0, 50), ("leading", 1, 25), ("trailing", 0, 10)):
        ours = DPMSolverMultistepScheduler(timestep_spacing=spacing, steps_offset=off)
        ref = sampler_ref.DPMSolverMultistepRef(timestep_spacing=spacing, steps_offset=off)
        assert ours.set_timesteps(n).tolist() == ref.set_timesteps(n).tolist()
        assert np.allclose(ours.sigmas, ref.sigmas, rtol=1e-6, atol=0)


@pytest.mark.parametrize("pred", ["epsilon", "v_prediction"])
@pytest.mark.parametrize("spacing,n", [("linspace", 25), ("leading", 50), ("trailing", 6), ("linspace", 2),
                                       ("linspace", 1)])
def test_step_coefficients_reproduce_reference_update(pred, spacing, n):
    """x <- c_x x + c_d0 m0 + c_d1 (m0 - m_prev) with the host coefficients equals DPMSolverMultistepScheduler.step
    (first order on the first and last step, midpoint second order in between), in float64."""
    from oracle import sampler_ref
    from textboost_b200.pipeline import DPMSolverMultistepScheduler
    ours = DPMSolverMultistepScheduler(prediction_type=pred, timestep_spacing=spacing)
    ref = sampler_ref.DPMSolverMultistepRef(prediction_type=pred, timestep_spacing=spacing)
    ours.set_timesteps(n)
    ref.set_timesteps(n)
    g = torch.Generator().manual_seed(n)
    x = torch.randn(2, 4, 8, 8, generator=g, dtype=torch.float64) * ref.sigmas[0]
    xr = x.clone()
    m_prev = None
    for i in range(n):
        e = torch.randn(x.shape, generator=g, dtype=torch.float64)
        k = ours.step_coefficients(i)
        m0 = k["alpha_i"] * x - k["sigma_i"] * e if pred == "v_prediction" else (x - k["sigma_i"] * e) / k["alpha_i"]
        assert (k["c_d1"] != 0.0) == (0 < i < n - 1)
        x = k["c_x"] * x + k["c_d0"] * m0 + (k["c_d1"] * (m0 - m_prev) if k["c_d1"] else 0.0)
        m_prev = m0
        xr = ref.step(e, xr)
        assert torch.allclose(x, xr, rtol=1e-9, atol=1e-9), i
    last = ours.step_coefficients(n - 1)
    assert last["c_x"] == 0.0 and abs(last["c_d0"] - 1.0) < 1e-12  # final sigma zero: the last step returns m0


def test_dpm_solver_recovers_fixed_point():
    """Size-independent property: when the model returns the exact noise of a fixed x0, DPM-Solver++ lands on x0."""
    from textboost_b200.pipeline import DPMSolverMultistepScheduler
    s = DPMSolverMultistepScheduler()
    s.set_timesteps(20)
    x0 = torch.randn(3, 4, 8, 8, dtype=torch.float64)
    a0 = 1 / math.sqrt(s.sigmas[0] ** 2 + 1)
    x = a0 * x0 + s.sigmas[0] * a0 * torch.randn(x0.shape, dtype=torch.float64)
    m_prev = None
    for i in range(20):
        k = s.step_coefficients(i)
        e = (x - k["alpha_i"] * x0) / k["sigma_i"]
        m0 = (x - k["sigma_i"] * e) / k["alpha_i"]
        x = k["c_x"] * x + k["c_d0"] * m0 + (k["c_d1"] * (m0 - m_prev) if k["c_d1"] else 0.0)
        m_prev = m0
    assert (x - x0).abs().max() < 1e-12


def test_scheduler_from_config_drops_foreign_keys(tmp_path):
    from textboost_b200.pipeline import DPMSolverMultistepScheduler
    pndm = {"_class_name": "PNDMScheduler", "_diffusers_version": "0.6.0", "beta_end": 0.012,
            "beta_schedule": "scaled_linear", "beta_start": 0.00085, "num_train_timesteps": 1000,
            "set_alpha_to_one": False, "skip_prk_steps": True, "steps_offset": 1, "trained_betas": None,
            "clip_sample": False}
    s = DPMSolverMultistepScheduler.from_config(pndm)
    assert s.config.timestep_spacing == "linspace" and s.config.steps_offset == 1
    assert s.set_timesteps(25)[0] == 999  # steps_offset only matters for "leading"
    (tmp_path / "scheduler").mkdir()
    with open(tmp_path / "scheduler" / "scheduler_config.json", "w") as f:
        json.dump({**pndm, "timestep_spacing": "leading", "prediction_type": "v_prediction"}, f)
    s2 = DPMSolverMultistepScheduler.from_pretrained(str(tmp_path))
    assert s2.config.prediction_type == "v_prediction" and s2.set_timesteps(25)[0] == 951
    s3 = DPMSolverMultistepScheduler.from_config(s2.config)
    assert s3.config.timestep_spacing == "leading"
    with pytest.raises(NotImplementedError):
        DPMSolverMultistepScheduler(beta_schedule="linear")
    with pytest.raises(NotImplementedError):
        DPMSolverMultistepScheduler(algorithm_type="sde-dpmsolver++")


class _FakeUNetEngine:
    """A smooth, batch-independent stand-in for the denoiser (the real one is checked on the GPU): what matters here
    is that the loop feeds it the doubled batch, the right timesteps and conditioning, in the right order."""

    def __init__(self):
        self.calls = []

    def __call__(self, x, t, ehs):
        return self.fn(x.float(), t, ehs.float())

    @staticmethod
    def fn(x, t, ehs):
        c = ehs.mean(dim=(1, 2)).view(-1, 1, 1, 1)
        return torch.tanh(0.3 * x + c) * (1.0 + t.view(-1, 1, 1, 1).float() / 1000.0)

    def forward(self, x, t, ehs, save_for_backward=False):
        assert x.dtype == torch.float16 and t.dtype == torch.int64 and ehs.dtype == torch.float16
        assert not save_for_backward
        self.calls.append((x.shape[0], int(t[0])))
        return self.fn(x.float(), t, ehs.float()).half()


@pytest.mark.parametrize("guidance,steps", [(7.5, 12), (1.0, 5)])
def test_denoise_loop_matches_oracle_sampler(monkeypatch, guidance, steps):
    from oracle import sampler_ref
    from textboost_b200 import pipeline
    ops_standin.install(monkeypatch)
    eng = _FakeUNetEngine()
    unet = SimpleNamespace(engine=eng, config=SimpleNamespace(sample_size=8, in_channels=4))
    vae = SimpleNamespace(config={"block_out_channels": (1, 2, 3, 4)})
    pipe = pipeline.StableDiffusionPipeline(vae, None, None, unet, pipeline.DPMSolverMultistepScheduler())
    pipe.use_cuda_graph = False
    g = torch.Generator().manual_seed(3)
    cond = (torch.randn(3, 7, 16, generator=g) * 0.5).half()
    uncond = (torch.randn(3, 7, 16, generator=g) * 0.5).half()
    lat = pipe.prepare_latents(3, 64, 64, "cpu", generator=torch.Generator().manual_seed(5))
    assert lat.shape == (3, 4, 8, 8) and lat.dtype == torch.float32
    x = pipe.denoise(lat.clone(), cond, uncond, steps, guidance)
    ref = sampler_ref.sample_latents(lambda a, t, e: eng.fn(a, t, e), cond.float(), uncond.float(), lat.clone(),
                                     sampler_ref.DPMSolverMultistepRef(), steps, guidance)
    assert [c[0] for c in eng.calls] == [6 if guidance > 1 else 3] * steps
    assert [c[1] for c in eng.calls] == pipe.scheduler.timesteps.tolist()
    err = ((x - ref).norm() / ref.norm()).item()
    assert err < 5e-3, err  # fp16 model input / output per step vs the fp32 oracle loop
    # a list of generators draws image i from generator i (inference.py:91-93)
    gens = [torch.Generator().manual_seed(s) for s in (0, 1, 2)]
    per = pipe.prepare_latents(3, 64, 64, "cpu", generator=gens)
    assert torch.equal(per[1:2], torch.randn((1, 4, 8, 8), generator=torch.Generator().manual_seed(1)))
    with pytest.raises(ValueError):
        pipe.prepare_latents(3, 64, 64, "cpu", generator=gens[:2])
    with pytest.raises(ValueError):
        pipe.prepare_latents(3, 64, 64, "cpu", latents=torch.zeros(1, 4, 8, 8))


def test_inference_cli_flags_match_reference_table(tmp_path):
    """inference.py keeps the reference's positional argument and flags (tests/golden/inference_flags.json, extracted
    from /root/reference/inference.py's AST); --num_inference_steps / --guidance_scale are additions."""
    import argparse
    import os
    import inference as I
    with open(os.path.join(os.path.dirname(__file__), "golden", "inference_flags.json")) as f:
        gold = json.load(f)["flags"]
    parser_actions = {}
    real = argparse.ArgumentParser.add_argument

    def spy(self, *names, **kw):
        parser_actions[names[0]] = kw
        return real(self, *names, **kw)

    argparse.ArgumentParser.add_argument = spy
    try:
        a = I.parse_args(["some/dir"])
    finally:
        argparse.ArgumentParser.add_argument = real
    for name, spec in gold.items():
        kw = parser_actions[name]
        for k, v in spec.items():
            got = kw.get(k)
            assert (got.__name__ if k == "type" else got) == v, (name, k, got, v)
    assert set(parser_actions) - set(gold) - {"-h"} == {"--num_inference_steps", "--guidance_scale"}
    assert a.path == "some/dir" and a.seeds == [0, 1, 2, 3] and a.prompt == "photo of a <dog> dog"
    assert a.model == "stabilityai/stable-diffusion-2-1-base"  # short names map to hub ids (inference.py:15-20, 41-42)
    d = tmp_path / "sd21base"
    d.mkdir()
    assert I.parse_args(["x", "--model", str(d)]).model == str(d)  # ... unless they are local directories
    g = I.make_image_grid([__import__("PIL.Image").Image.new("RGB", (8, 6))] * 6, 2, 3)
    assert g.size == (24, 12)


@pytest.mark.parametrize("pred", ["epsilon", "v_prediction"])
@pytest.mark.parametrize("spacing,off,n,var", [("leading", 1, 25, "fixed_small"), ("linspace", 0, 10, "fixed_small"),
                                               ("trailing", 0, 7, "fixed_large"), ("leading", 0, 1000, "fixed_small")])
def test_ddpm_step_coefficients_reproduce_reference_update(pred, spacing, off, n, var):
    """--validation_scheduler DDPMScheduler: x <- c_x x + c_d0 x0 + noise_std z with the host coefficients of the mirror
    equals DDPMScheduler.step as restated in oracle/sampler_ref.DDPMRef (posterior mean + fixed variance), float64;
    the timestep grids match, and the last step (t = 0 on the leading / linspace grids without offset) adds no noise."""
    from oracle import sampler_ref
    from textboost_b200.pipeline import DDPMScheduler
    ours = DDPMScheduler(prediction_type=pred, timestep_spacing=spacing, steps_offset=off, variance_type=var)
    ref = sampler_ref.DDPMRef(prediction_type=pred, timestep_spacing=spacing, steps_offset=off, variance_type=var)
    assert ours.set_timesteps(n).tolist() == ref.set_timesteps(n).tolist()
    g = torch.Generator().manual_seed(n)
    x = torch.randn(2, 4, 8, 8, generator=g, dtype=torch.float64)
    xr = x.clone()
    steps = range(n) if n <= 25 else list(range(3)) + [n - 2, n - 1]
    for i in steps:
        ref.step_index = i
        e = torch.randn(x.shape, generator=g, dtype=torch.float64)
        k = ours.step_coefficients(i)
        assert k["c_d1"] == 0.0
        x0 = k["alpha_i"] * x - k["sigma_i"] * e if pred == "v_prediction" else (x - k["sigma_i"] * e) / k["alpha_i"]
        gz = torch.Generator().manual_seed(100 + i)
        z = torch.randn(x.shape, generator=gz, dtype=torch.float32).double()
        x = k["c_x"] * x + k["c_d0"] * x0 + k["noise_std"] * z
        xr = ref.step(e, xr, torch.Generator().manual_seed(100 + i))
        assert torch.allclose(x, xr, rtol=1e-9, atol=1e-9), i
        assert (k["noise_std"] == 0.0) == (int(ours.timesteps[i]) == 0)
    # the reference maps learned variances onto fixed_small (train_textboost.py:488-489); from_config keeps the SD keys
    s = DDPMScheduler.from_config({"num_train_timesteps": 1000, "beta_schedule": "scaled_linear", "beta_start": 0.00085,
                                   "beta_end": 0.012, "variance_type": "learned_range", "clip_sample": False,
                                   "skip_prk_steps": True, "set_alpha_to_one": False, "steps_offset": 1,
                                   "timestep_spacing": "leading", "_class_name": "PNDMScheduler"})
    assert s.config.variance_type == "fixed_small" and s.config.steps_offset == 1


def test_denoise_loop_with_ddpm_scheduler_matches_oracle(monkeypatch):
    """The pipeline's loop with the DDPM mirror against the oracle loop: same generator, one noise draw per step, the
    next UNet input written after the noise."""
    from oracle import sampler_ref
    from textboost_b200 import pipeline
    ops_standin.install(monkeypatch)
    eng = _FakeUNetEngine()
    unet = SimpleNamespace(engine=eng, config=SimpleNamespace(sample_size=8, in_channels=4))
    vae = SimpleNamespace(config={"block_out_channels": (1, 2, 3, 4)})
    pipe = pipeline.StableDiffusionPipeline(vae, None, None, unet,
                                            pipeline.DDPMScheduler(timestep_spacing="leading", steps_offset=1))
    pipe.use_cuda_graph = False
    g = torch.Generator().manual_seed(3)
    cond = (torch.randn(3, 7, 16, generator=g) * 0.5).half()
    uncond = (torch.randn(3, 7, 16, generator=g) * 0.5).half()
    lat = pipe.prepare_latents(3, 64, 64, "cpu", generator=torch.Generator().manual_seed(5))
    x = pipe.denoise(lat.clone(), cond, uncond, 9, 7.5, generator=torch.Generator().manual_seed(11))
    ref = sampler_ref.sample_latents(lambda a, t, e: eng.fn(a, t, e), cond.float(), uncond.float(), lat.clone(),
                                     sampler_ref.DDPMRef(), 9, 7.5, generator=torch.Generator().manual_seed(11))
    assert [c[1] for c in eng.calls] == pipe.scheduler.timesteps.tolist() == [889, 778, 667, 556, 445, 334, 223, 112, 1]
    err = ((x - ref).norm() / ref.norm()).item()
    assert err < 5e-3, err
