"""oracle (test infrastructure, parity UNPINNED: diffusers is not installable here): the validation / inference sampler
of the reference, i.e. diffusers 0.29 ``StableDiffusionPipeline.__call__`` with ``DPMSolverMultistepScheduler`` at its
defaults, as reached from

    /root/reference/train_textboost.py:453-531   log_validation: pipeline(prompt, num_images_per_prompt=n,
                                                 num_inference_steps=25, generator=g)      (--validation_scheduler)
    /root/reference/inference.py:84-105          pipeline(prompt, num_images_per_prompt=len(seeds), generator=[...])

restated in plain numpy / PyTorch.  Scheduler defaults restated: solver_order 2, algorithm_type "dpmsolver++",
solver_type "midpoint", lower_order_final True, final_sigmas_type "zero", no Karras sigmas, no thresholding,
timestep_spacing from the checkpoint's scheduler_config.json ("linspace" when absent, "leading" + steps_offset for the
SD-2.x configs).  Pipeline defaults: guidance_scale 7.5, negative prompt "", 50 steps, eta unused.
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np
import torch


class DPMSolverMultistepRef:
    """The slice of DPMSolverMultistepScheduler the pipeline touches: set_timesteps / step / init_noise_sigma."""

    init_noise_sigma = 1.0

    def __init__(self, alphas_cumprod=None, prediction_type="epsilon", timestep_spacing="linspace", steps_offset=0,
                 solver_order=2):
        if alphas_cumprod is None:
            from . import ddpm_ref
            alphas_cumprod = ddpm_ref.alphas_cumprod().numpy()
        self.acp = np.asarray(alphas_cumprod, dtype=np.float64)
        self.T = len(self.acp)
        self.prediction_type = prediction_type
        self.timestep_spacing, self.steps_offset, self.order = timestep_spacing, steps_offset, solver_order

    def set_timesteps(self, n: int):
        T = self.T
        if self.timestep_spacing == "linspace":
            ts = np.linspace(0, T - 1, n + 1).round()[::-1][:-1].copy().astype(np.int64)
        elif self.timestep_spacing == "leading":
            ratio = T // (n + 1)
            ts = (np.arange(0, n + 1) * ratio).round()[::-1][:-1].copy().astype(np.int64) + self.steps_offset
        elif self.timestep_spacing == "trailing":
            ts = (np.arange(T, 0, -T / n).round() - 1).astype(np.int64)
        else:
            raise ValueError(self.timestep_spacing)
        sig_all = ((1 - self.acp) / self.acp) ** 0.5
        sig = np.interp(ts, np.arange(0, T), sig_all)
        self.sigmas = np.concatenate([sig, [0.0]]).astype(np.float32).astype(np.float64)  # final_sigmas_type "zero"
        self.timesteps = ts
        self.model_outputs: List[Optional[torch.Tensor]] = [None] * self.order
        self.lower_order_nums = 0
        self.step_index = 0
        return ts

    @staticmethod
    def _alpha_sigma(sigma):
        alpha_t = 1.0 / math.sqrt(sigma ** 2 + 1.0)
        return alpha_t, sigma * alpha_t

    @staticmethod
    def _lambda(alpha_t, sigma_t):
        return math.log(alpha_t) - (math.log(sigma_t) if sigma_t > 0 else -math.inf)

    def step(self, model_output: torch.Tensor, sample: torch.Tensor) -> torch.Tensor:
        i, n = self.step_index, len(self.timesteps)
        lower_order_final = i == n - 1  # final sigma is zero: the last step is always first order
        # (lower_order_second only demotes a third-order solver; at solver_order 2 it changes nothing)
        a_i, s_i = self._alpha_sigma(self.sigmas[i])
        if self.prediction_type == "epsilon":
            x0 = (sample - s_i * model_output) / a_i
        elif self.prediction_type == "v_prediction":
            x0 = a_i * sample - s_i * model_output
        else:
            raise ValueError(self.prediction_type)
        self.model_outputs = self.model_outputs[1:] + [x0]
        a_t, s_t = self._alpha_sigma(self.sigmas[i + 1])
        lam_t, lam_s0 = self._lambda(a_t, s_t), self._lambda(a_i, s_i)
        h = lam_t - lam_s0
        em1 = math.exp(-h) - 1.0
        if self.order == 1 or self.lower_order_nums < 1 or lower_order_final:
            prev = (s_t / s_i) * sample - (a_t * em1) * x0
        else:
            a_p, s_p = self._alpha_sigma(self.sigmas[i - 1])
            h0 = lam_s0 - self._lambda(a_p, s_p)
            r0 = h0 / h
            m0, m1 = self.model_outputs[-1], self.model_outputs[-2]
            d1 = (1.0 / r0) * (m0 - m1)
            prev = (s_t / s_i) * sample - (a_t * em1) * m0 - 0.5 * (a_t * em1) * d1
        if self.lower_order_nums < self.order:
            self.lower_order_nums += 1
        self.step_index += 1
        return prev


class DDPMRef:
    """diffusers 0.29 ``DDPMScheduler`` as ``--validation_scheduler DDPMScheduler`` uses it
    (/root/reference/train_textboost.py:341-346, 483-495): set_timesteps over strided timesteps, step = posterior mean
    of q(x_{t-1} | x_t, x0) (formula (7) of the DDPM paper) plus sqrt(variance) * noise for t > 0; variance_type
    "fixed_small" (learned variances are mapped onto it, :488-489) or "fixed_large"; clip_sample False, no thresholding
    (the SD scheduler configs)."""

    init_noise_sigma = 1.0

    def __init__(self, alphas_cumprod=None, prediction_type="epsilon", timestep_spacing="leading", steps_offset=1,
                 variance_type="fixed_small"):
        if alphas_cumprod is None:
            from . import ddpm_ref
            alphas_cumprod = ddpm_ref.alphas_cumprod().numpy()
        self.acp = np.asarray(alphas_cumprod, dtype=np.float64)
        self.T = len(self.acp)
        self.prediction_type, self.variance_type = prediction_type, variance_type
        self.timestep_spacing, self.steps_offset = timestep_spacing, steps_offset

    def set_timesteps(self, n: int):
        T = self.T
        if self.timestep_spacing == "linspace":
            ts = np.linspace(0, T - 1, n).round()[::-1].copy().astype(np.int64)
        elif self.timestep_spacing == "leading":
            ts = (np.arange(0, n) * (T // n)).round()[::-1].copy().astype(np.int64) + self.steps_offset
        elif self.timestep_spacing == "trailing":
            ts = np.round(np.arange(T, 0, -T / n)).astype(np.int64) - 1
        else:
            raise ValueError(self.timestep_spacing)
        self.timesteps, self.n, self.step_index = ts, n, 0
        return ts

    def step(self, model_output: torch.Tensor, sample: torch.Tensor, generator=None) -> torch.Tensor:
        t = int(self.timesteps[self.step_index])
        prev_t = t - self.T // self.n
        a_t = self.acp[t]
        a_prev = self.acp[prev_t] if prev_t >= 0 else 1.0
        b_t, b_prev = 1.0 - a_t, 1.0 - a_prev
        cur_alpha = a_t / a_prev
        cur_beta = 1.0 - cur_alpha
        if self.prediction_type == "epsilon":
            x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
        elif self.prediction_type == "v_prediction":
            x0 = a_t ** 0.5 * sample - b_t ** 0.5 * model_output
        else:
            raise ValueError(self.prediction_type)
        prev = (a_prev ** 0.5 * cur_beta / b_t) * x0 + (cur_alpha ** 0.5 * b_prev / b_t) * sample
        if t > 0:
            var = max(b_prev / b_t * cur_beta, 1e-20) if self.variance_type == "fixed_small" else cur_beta
            if isinstance(generator, (list, tuple)):
                z = torch.cat([torch.randn((1,) + tuple(sample.shape[1:]), generator=g, device=sample.device,
                                           dtype=torch.float32) for g in generator])
            else:
                z = torch.randn(sample.shape, generator=generator, device=sample.device, dtype=torch.float32)
            prev = prev + var ** 0.5 * z.to(sample.dtype)
        self.step_index += 1
        return prev


@torch.no_grad()
def sample_latents(unet, cond, uncond, latents, scheduler, num_inference_steps=50, guidance_scale=7.5,
                   generator=None):
    """The denoising loop of StableDiffusionPipeline.__call__: `unet(x, t, ehs)` on the [uncond | cond] doubled batch,
    classifier-free guidance, scheduler.step.  cond / uncond [B,L,D]; latents [B,4,h,w] unit Gaussian."""
    ts = scheduler.set_timesteps(num_inference_steps)
    x = latents * scheduler.init_noise_sigma
    ehs = torch.cat([uncond, cond])
    for t in ts:
        tt = torch.full((2 * x.shape[0],), int(t), dtype=torch.int64, device=x.device)
        eps = unet(torch.cat([x, x]), tt, ehs)
        e_u, e_c = eps.chunk(2)
        out = e_u + guidance_scale * (e_c - e_u)
        x = scheduler.step(out, x, generator) if isinstance(scheduler, DDPMRef) else scheduler.step(out, x)
    return x
