"""Validation / inference sampler on the B200 kernels: text -> images.

Mirror of the slice of diffusers the reference uses to LOOK at a trained model:

    /root/reference/train_textboost.py:453-531  log_validation: DiffusionPipeline.from_pretrained(path, vae=, tokenizer=,
        text_encoder=, unet=, safety_checker=None, ...), scheduler swapped for ``args.validation_scheduler``
        (DPMSolverMultistepScheduler) ``.from_config(pipeline.scheduler.config)``, then
        ``pipeline(prompt=..., num_images_per_prompt=n, num_inference_steps=25, generator=g).images``
    /root/reference/inference.py:84-105         the same pipeline with 50 steps and one generator per seed

What runs where:
  * text encoder (CLIP + LoRA + learned rows, null-embedding override): the training engine, no grad;
  * every denoising step: ONE UNet forward on the [uncond | cond] doubled batch (the training path's kernels, optionally
    replayed from one CUDA graph) + ONE fused launch (`tb_dpm_cfg_step`) for classifier-free guidance, the data
    prediction and the DPM-Solver++(2M) update, which also writes the next step's fp16 doubled input;
  * latents stay fp32 across steps (diffusers keeps them in the pipeline dtype, fp16: ours is the tighter of the two);
  * VAE decoder engine + `tb_image_u8` for the uint8 image.
The scheduler's per-step scalars (sigma schedule, exponential-integrator coefficients) are host float64 arithmetic, as
in diffusers; parity of this module is UNPINNED (diffusers is not installable here) and checked against
oracle/sampler_ref.py.
"""
from __future__ import annotations

import json
import math
import os
from types import SimpleNamespace
from typing import List, Optional, Sequence, Union

import numpy as np
import torch

from . import ops

from .precision import POLICY
F32 = torch.float32


class DPMSolverMultistepScheduler:
    """dpmsolver++ / midpoint / order 2 / final sigma zero (the diffusers defaults the reference runs with)."""

    init_noise_sigma = 1.0
    order = 1

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                 prediction_type="epsilon", timestep_spacing="linspace", steps_offset=0, solver_order=2,
                 algorithm_type="dpmsolver++", solver_type="midpoint", **_other_scheduler_keys):
        if beta_schedule != "scaled_linear":
            raise NotImplementedError(f"beta_schedule {beta_schedule!r}: SD checkpoints use scaled_linear")
        if algorithm_type != "dpmsolver++" or solver_type != "midpoint" or solver_order not in (1, 2):
            raise NotImplementedError("only the default dpmsolver++ / midpoint solver of order <= 2 is built")
        if prediction_type not in ("epsilon", "v_prediction"):
            raise NotImplementedError(f"prediction_type {prediction_type!r}")
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, beta_start=beta_start,
                                      beta_end=beta_end, beta_schedule=beta_schedule,
                                      prediction_type=prediction_type, timestep_spacing=timestep_spacing,
                                      steps_offset=steps_offset, solver_order=solver_order)
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=F32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.timesteps: Optional[np.ndarray] = None
        self.sigmas: Optional[np.ndarray] = None

    @classmethod
    def from_config(cls, config, **overrides):
        cfg = dict(config) if isinstance(config, dict) else dict(vars(config))
        cfg.update(overrides)
        cfg = {k: v for k, v in cfg.items() if not k.startswith("_")}
        # keys of other scheduler classes (PNDM's skip_prk_steps, set_alpha_to_one, ...) are dropped, as from_config does
        keep = ("num_train_timesteps", "beta_start", "beta_end", "beta_schedule", "prediction_type",
                "timestep_spacing", "steps_offset", "solver_order", "algorithm_type", "solver_type")
        return cls(**{k: cfg[k] for k in keep if k in cfg})

    @classmethod
    def from_pretrained(cls, path, subfolder="scheduler", **overrides):
        with open(os.path.join(path, subfolder, "scheduler_config.json")) as f:
            return cls.from_config(json.load(f), **overrides)

    def set_timesteps(self, num_inference_steps: int, device=None):
        c, n = self.config, int(num_inference_steps)
        T = c.num_train_timesteps
        if c.timestep_spacing == "linspace":
            ts = np.linspace(0, T - 1, n + 1).round()[::-1][:-1]
        elif c.timestep_spacing == "leading":
            ts = (np.arange(0, n + 1) * (T // (n + 1))).round()[::-1][:-1] + c.steps_offset
        elif c.timestep_spacing == "trailing":
            ts = np.arange(T, 0, -T / n).round() - 1
        else:
            raise ValueError(f"timestep_spacing {c.timestep_spacing!r}")
        ts = ts.copy().astype(np.int64)
        acp = self.alphas_cumprod.double().numpy()
        table = np.sqrt((1.0 - acp) / acp)
        sig = np.interp(ts, np.arange(T), table)
        self.sigmas = np.append(sig, 0.0).astype(np.float32).astype(np.float64)
        self.timesteps = ts
        return ts

    def step_coefficients(self, i: int) -> dict:
        """Host scalars of denoising step i for `tb_dpm_cfg_step`: (alpha_i, sigma_i) turn the model output into the
        data prediction m0; x <- c_x x + c_d0 m0 + c_d1 (m0 - m_prev).  First order on the first and the last step."""
        sig, n = self.sigmas, len(self.timesteps)

        def alpha_sigma(s):
            a = 1.0 / math.sqrt(s * s + 1.0)
            return a, s * a

        def lam(s):
            a, sg = alpha_sigma(s)
            return math.log(a) - math.log(sg) if sg > 0 else math.inf

        a_i, s_i = alpha_sigma(sig[i])
        a_t, s_t = alpha_sigma(sig[i + 1])
        h = lam(sig[i + 1]) - lam(sig[i])
        c_d0 = -a_t * math.expm1(-h)
        c_d1 = 0.0
        if self.config.solver_order == 2 and 0 < i < n - 1:
            r0 = (lam(sig[i]) - lam(sig[i - 1])) / h
            c_d1 = 0.5 * c_d0 / r0
        return dict(alpha_i=a_i, sigma_i=s_i, c_x=s_t / s_i, c_d0=c_d0, c_d1=c_d1)


class DDPMScheduler(DPMSolverMultistepScheduler):
    """``--validation_scheduler DDPMScheduler`` (train_textboost.py:341-346, 493-495): ancestral DDPM sampling over
    ``num_inference_steps`` strided timesteps, diffusers 0.29 ``DDPMScheduler`` at the settings an SD scheduler config
    gives it (variance_type "fixed_small" — the reference maps learned variances onto it, :485-491 —, no sample clipping,
    no thresholding).  Shares the config handling of the DPM-Solver mirror; the update is
    ``x <- c_xt x + c_x0 x0 + sqrt(var) z`` with x0 the data prediction: the first two terms are one call of the fused
    CFG + update kernel (``tb_dpm_cfg_step`` with no second-order term), the noise is drawn from the pipeline's generator."""

    def __init__(self, *a, variance_type="fixed_small", clip_sample=False, **kw):
        kw.pop("solver_order", None), kw.pop("algorithm_type", None), kw.pop("solver_type", None)
        super().__init__(*a, **kw)
        if variance_type in ("learned", "learned_range"):
            variance_type = "fixed_small"  # train_textboost.py:488-489
        if variance_type not in ("fixed_small", "fixed_large"):
            raise NotImplementedError(f"DDPMScheduler variance_type {variance_type!r}")
        if clip_sample:
            raise NotImplementedError("DDPMScheduler clip_sample=True (SD scheduler configs set it to false)")
        self.config.variance_type = variance_type

    @classmethod
    def from_config(cls, config, **overrides):
        cfg = dict(config) if isinstance(config, dict) else dict(vars(config))
        cfg.update(overrides)
        keep = ("num_train_timesteps", "beta_start", "beta_end", "beta_schedule", "prediction_type",
                "timestep_spacing", "steps_offset", "variance_type", "clip_sample")
        return cls(**{k: cfg[k] for k in keep if k in cfg})

    def set_timesteps(self, num_inference_steps: int, device=None):
        c, n = self.config, int(num_inference_steps)
        T = c.num_train_timesteps
        if c.timestep_spacing == "linspace":
            ts = np.linspace(0, T - 1, n).round()[::-1]
        elif c.timestep_spacing == "leading":
            ts = (np.arange(0, n) * (T // n)).round()[::-1] + c.steps_offset
        elif c.timestep_spacing == "trailing":
            ts = np.round(np.arange(T, 0, -T / n)) - 1
        else:
            raise ValueError(f"timestep_spacing {c.timestep_spacing!r}")
        self.timesteps = ts.copy().astype(np.int64)
        self.num_inference_steps = n
        self.sigmas = None
        return self.timesteps

    def step_coefficients(self, i: int) -> dict:
        """Host scalars of step i: (alpha_i, sigma_i) give the data prediction x0 = (x - sigma eps) / alpha (or
        alpha x - sigma v); x <- c_x x + c_d0 x0 + noise_std z  (z omitted at t = 0).  DDPMScheduler.step / _get_variance."""
        t = int(self.timesteps[i])
        prev_t = t - self.config.num_train_timesteps // self.num_inference_steps
        acp = self.alphas_cumprod.double()
        a_t = float(acp[t])
        a_prev = float(acp[prev_t]) if prev_t >= 0 else 1.0
        b_t, b_prev = 1.0 - a_t, 1.0 - a_prev
        cur_alpha = a_t / a_prev
        cur_beta = 1.0 - cur_alpha
        var = max(b_prev / b_t * cur_beta, 1e-20) if self.config.variance_type == "fixed_small" else cur_beta
        return dict(alpha_i=math.sqrt(a_t), sigma_i=math.sqrt(b_t), c_x=math.sqrt(cur_alpha) * b_prev / b_t,
                    c_d0=math.sqrt(a_prev) * cur_beta / b_t, c_d1=0.0, noise_std=math.sqrt(var) if t > 0 else 0.0)


class StableDiffusionPipelineOutput:
    def __init__(self, images):
        self.images = images


class StableDiffusionPipeline:
    """``pipeline(prompt, num_images_per_prompt=..., num_inference_steps=..., generator=...).images``."""

    def __init__(self, vae, text_encoder, tokenizer, unet, scheduler, safety_checker=None, **_unused):
        self.vae, self.text_encoder, self.tokenizer, self.unet, self.scheduler = vae, text_encoder, tokenizer, unet, \
            scheduler
        self.safety_checker = None
        self.vae_scale_factor = 2 ** (len(vae.config["block_out_channels"]) - 1)
        self.use_cuda_graph = True
        self.last_latents: Optional[torch.Tensor] = None  # final latents of the last call (tests, debugging)

    # ------------------------------------------------------------------ construction (DiffusionPipeline surface)
    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, vae=None, text_encoder=None, tokenizer=None, unet=None,
                        scheduler=None, safety_checker=None, revision=None, variant=None, torch_dtype=None,
                        use_safetensors=None, **_unused):
        path = pretrained_model_name_or_path
        if vae is None:
            from .vae import AutoencoderKL
            vae = AutoencoderKL.from_pretrained(path, subfolder="vae", revision=revision, variant=variant)
        if text_encoder is None:  # a stock checkpoint carries a plain CLIPTextModel
            from .text_encoder import CLIPTextModel
            text_encoder = CLIPTextModel.from_pretrained(path, subfolder="text_encoder", revision=revision,
                                                         variant=variant)
        if unet is None:
            from .unet_model import UNet2DConditionModel
            unet = UNet2DConditionModel.from_pretrained(path, subfolder="unet", revision=revision, variant=variant)
        if tokenizer is None:
            tokenizer = load_tokenizer(path)
        if scheduler is None:
            scheduler = DPMSolverMultistepScheduler.from_pretrained(path)
        return cls(vae, text_encoder, tokenizer, unet, scheduler)

    def to(self, device=None, dtype=None):
        if device is not None:
            self.vae.to(device)
            self.text_encoder.to(device)
            self.unet.to(device, dtype=POLICY.act)
        return self

    def set_progress_bar_config(self, **_kw):
        return None

    def load_textual_inversion(self, path, token=None, **_unused):
        """diffusers TextualInversionLoaderMixin on a ``{token: embedding}`` file as train_textboost.py:1188-1209 writes
        them (``[D]`` placeholder rows, ``[1, D]`` augmentation rows; ``[n, D]`` becomes token, token_1, ...): the
        tokens join the tokenizer, the rows the embedding matrix.  Must precede ``.to('cuda')``."""
        sd = torch.load(path, map_location="cpu", weights_only=True)
        tokens, rows = [], []
        for name, emb in sd.items():
            name = token or name
            if emb.dim() > 1 and emb.shape[0] > 1:
                tokens += [name] + [f"{name}_{i}" for i in range(1, emb.shape[0])]
                rows += list(emb)
            else:
                tokens.append(name)
                rows.append(emb[0] if emb.dim() > 1 else emb)
        vocab = getattr(self.tokenizer, "get_vocab", lambda: getattr(self.tokenizer, "added", {}))()
        for t in tokens:
            if t in vocab:
                raise ValueError(f"Token {t} already in tokenizer vocabulary. Please choose a different token name or "
                                 "remove it from the embedding file.")
        self.tokenizer.add_tokens(tokens)
        ids = self.tokenizer.convert_tokens_to_ids(tokens)
        emb_layer = self.text_encoder.resize_token_embeddings(len(self.tokenizer))
        for i, row in zip(ids, rows):
            emb_layer.weight.data[i] = row.to(emb_layer.weight.dtype)

    # ------------------------------------------------------------------ pieces
    def _tokenize(self, prompts: Sequence[str]) -> torch.Tensor:
        tok = self.tokenizer
        rows = [tok(p, truncation=True, padding="max_length", max_length=tok.model_max_length,
                    return_tensors="pt").input_ids for p in prompts]
        return torch.cat(rows, dim=0)

    def encode_prompt(self, prompt, device, num_images_per_prompt=1, negative_prompt=None):
        """-> (cond, uncond) fp16 [n_prompts * num_images, L, D]; images of one prompt are contiguous."""
        prompts = [prompt] if isinstance(prompt, str) else list(prompt)
        if negative_prompt is None:
            negatives = [""] * len(prompts)
        else:
            negatives = [negative_prompt] * len(prompts) if isinstance(negative_prompt, str) else list(negative_prompt)
        if len(negatives) != len(prompts):
            raise ValueError("`negative_prompt` must have the same batch size as `prompt`")
        with torch.no_grad():
            cond = self.text_encoder(self._tokenize(prompts).to(device))[0]
            uncond = self.text_encoder(self._tokenize(negatives).to(device))[0]
        rep = lambda e: e.to(POLICY.act).repeat_interleave(num_images_per_prompt, dim=0).contiguous()  # noqa: E731
        return rep(cond), rep(uncond)

    def prepare_latents(self, n, height, width, device, generator=None, latents=None):
        shape = (n, self.unet.config.in_channels, height // self.vae_scale_factor, width // self.vae_scale_factor)
        if latents is not None:
            if tuple(latents.shape) != shape:
                raise ValueError(f"Unexpected latents shape, got {tuple(latents.shape)}, expected {shape}")
            x = latents.to(device=device, dtype=F32)
        elif isinstance(generator, (list, tuple)):
            if len(generator) != n:
                raise ValueError(f"You have passed a list of generators of length {len(generator)}, but requested an "
                                 f"effective batch size of {n}.")
            x = torch.cat([torch.randn((1,) + shape[1:], generator=g, device=device, dtype=F32) for g in generator])
        else:
            x = torch.randn(shape, generator=generator, device=device, dtype=F32)
        return (x * self.scheduler.init_noise_sigma).contiguous()

    def denoise(self, x, cond, uncond, num_inference_steps, guidance_scale, generator=None):
        """The sampling loop on fp32 latents x [N,4,h,w] (in place); returns x.  `generator`: the ancestral noise of a
        DDPMScheduler (one draw per step, like scheduler.step(..., generator=generator) of the diffusers pipeline)."""
        sch, eng = self.scheduler, self.unet.engine
        ts = sch.set_timesteps(num_inference_steps)
        N = x.shape[0]
        cfg = guidance_scale > 1.0
        ehs = torch.cat([uncond, cond]) if cfg else cond
        nb = 2 * N if cfg else N
        unet_in = torch.empty((2 * N,) + tuple(x.shape[1:]), device=x.device, dtype=POLICY.act)
        unet_in[:N] = x
        unet_in[N:] = x
        tt = torch.empty(nb, device=x.device, dtype=torch.int64)
        m = [torch.empty_like(x), torch.empty_like(x)]
        v_pred = sch.config.prediction_type == "v_prediction"
        graph = eps_static = None
        for i, t in enumerate(ts):
            tt.fill_(int(t))
            if graph is not None:
                graph.replay()
                eps = eps_static
            else:
                eps = eng.forward(unet_in[:nb], tt, ehs, save_for_backward=False)
                if self.use_cuda_graph and i == 0 and len(ts) > 2:
                    # the first step ran eagerly (it also configures kernel attributes); every later UNet forward is
                    # a replay of one captured graph over the same static buffers
                    graph, eps_static = self._capture_forward(eng, unet_in[:nb], tt, ehs)
            if not cfg:
                eps = torch.cat([eps, eps])
            k = sch.step_coefficients(i)
            noise_std = k.get("noise_std", 0.0)
            last = i + 1 == len(ts)
            ops.dpm_cfg_step(x, eps.contiguous(), m[(i + 1) % 2] if k["c_d1"] != 0.0 else None, m[i % 2],
                             unet_in if not last and noise_std == 0.0 else None, guidance_scale if cfg else 1.0,
                             k["alpha_i"], k["sigma_i"], v_pred, k["c_x"], k["c_d0"], k["c_d1"])
            if noise_std != 0.0:  # ancestral step: the next UNet input is written after the noise has been added
                if isinstance(generator, (list, tuple)):
                    z = torch.cat([torch.randn((1,) + tuple(x.shape[1:]), generator=g, device=x.device, dtype=F32)
                                   for g in generator])
                else:
                    z = torch.randn(x.shape, generator=generator, device=x.device, dtype=F32)
                x.add_(z, alpha=noise_std)
                if not last:
                    x2 = x.view(N, -1)
                    ops.cast_f32_f16(x2, out=unet_in[:N].view(N, -1))
                    ops.cast_f32_f16(x2, out=unet_in[N:].view(N, -1))
        return x

    @staticmethod
    def _capture_forward(eng, unet_in, tt, ehs):
        # torch's process-wide capture stream, as TextBoostTrainer.capture: one split-K workspace is registered for it
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = eng.forward(unet_in, tt, ehs, save_for_backward=False)
        return graph, out

    # ------------------------------------------------------------------ the call
    @torch.no_grad()
    def __call__(self, prompt: Union[str, List[str]], height: Optional[int] = None, width: Optional[int] = None,
                 num_inference_steps: int = 50, guidance_scale: float = 7.5, negative_prompt=None,
                 num_images_per_prompt: int = 1, generator=None, latents=None, output_type: str = "pil",
                 return_dict: bool = True, **_ignored):
        dev = self.unet.device
        if dev.type != "cuda":
            raise RuntimeError("StableDiffusionPipeline: call .to('cuda') first (no CPU path)")
        size = self.unet.config.sample_size * self.vae_scale_factor
        height, width = height or size, width or size
        if height % 8 or width % 8:
            raise ValueError(f"`height` and `width` have to be divisible by 8 but are {height} and {width}.")
        cond, uncond = self.encode_prompt(prompt, dev, num_images_per_prompt, negative_prompt)
        x = self.prepare_latents(cond.shape[0], height, width, dev, generator, latents)
        x = self.denoise(x, cond, uncond, num_inference_steps, guidance_scale, generator)
        self.last_latents = x
        if output_type == "latent":
            images = x
        else:
            u8 = self.vae.decoder_engine.decode_u8(x)
            if output_type == "pt":
                images = u8
            elif output_type == "np":
                images = u8.cpu().numpy().astype(np.float32) / 255.0
            elif output_type == "pil":
                from PIL import Image
                images = [Image.fromarray(a) for a in u8.cpu().numpy()]
            else:
                raise ValueError(f"output_type {output_type!r}")
        return StableDiffusionPipelineOutput(images) if return_dict else (images, None)


DiffusionPipeline = StableDiffusionPipeline  # the name the reference imports (train_textboost.py:29, inference.py:6)


def load_tokenizer(path, subfolder="tokenizer"):
    """CLIP tokenizer of the checkpoint; the literal stand-in only for a synthetic checkpoint (marker file written by
    synthetic.write_pretrained).  Missing tokenizer files on a real checkpoint raise OSError."""
    from .synthetic import load_tokenizer as _load
    return _load(os.path.join(path, subfolder))
