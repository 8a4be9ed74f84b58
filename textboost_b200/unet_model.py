"""``UNet2DConditionModel`` — host-side mirror of the diffusers class on the TextBoost path, over the B200 UNet
engine (textboost_b200.unet.UNetEngine -> libtextboost_b200.so).

Surface kept from /root/reference/train_textboost.py:

  UNet2DConditionModel.from_pretrained(path, subfolder="unet", revision, variant)      :654-656
  unet.eval().requires_grad_(False)                                                    :696
  unet.to(device, dtype=weight_dtype) / unet.dtype / unet.config                       :810, :937
  unet(noisy_latents, timesteps, encoder_hidden_states).sample                         :1063-1067
  loss.backward() reaching encoder_hidden_states through the frozen UNet               :1108

The checkpoint layout is diffusers' (``config.json`` + ``diffusion_pytorch_model[.variant].safetensors``,
keys as in SURVEY.md Appendix A.4).  The UNet is frozen on this path (``--unet_params_to_train none``), so
``forward`` is a ``torch.autograd.Function`` whose backward is the hand-derived activation-backward (dgrad
only) and returns a gradient for ``encoder_hidden_states`` alone.  There is no CPU execution path.
"""
from __future__ import annotations

import json
import os
from types import SimpleNamespace
from typing import Dict, Optional

import torch

from .unet import UNetConfig, UNetEngine

from .precision import POLICY
F32 = torch.float32


class UNet2DConditionOutput:
    def __init__(self, sample):
        self.sample = sample

    def __getitem__(self, i):
        return (self.sample,)[i]


def _engine_config(raw: dict) -> UNetConfig:
    """diffusers ``unet/config.json`` -> UNetConfig; rejects what the SD-1.x / SD-2.x path does not use."""
    down = raw.get("down_block_types", ["CrossAttnDownBlock2D"] * 3 + ["DownBlock2D"])
    up = raw.get("up_block_types", ["UpBlock2D"] + ["CrossAttnUpBlock2D"] * 3)
    has_attn = tuple(t == "CrossAttnDownBlock2D" for t in down)
    if any(t not in ("CrossAttnDownBlock2D", "DownBlock2D") for t in down) or \
            [t == "CrossAttnUpBlock2D" for t in up] != list(reversed(has_attn)):
        raise NotImplementedError(f"unsupported block layout {down} / {up}")
    if raw.get("mid_block_type", "UNetMidBlock2DCrossAttn") != "UNetMidBlock2DCrossAttn":
        raise NotImplementedError("mid_block_type")
    for k, want in (("act_fn", "silu"), ("flip_sin_to_cos", True), ("freq_shift", 0), ("class_embed_type", None),
                    ("addition_embed_type", None), ("dual_cross_attention", False), ("only_cross_attention", False),
                    ("resnet_time_scale_shift", "default"), ("time_embedding_type", "positional")):
        if raw.get(k, want) != want:
            raise NotImplementedError(f"unet config {k}={raw[k]!r} is outside the SD-1.x/2.x TextBoost path")
    boc = tuple(raw.get("block_out_channels", (320, 640, 1280, 1280)))
    ahd = raw.get("attention_head_dim", 8)
    ahd = tuple(ahd) if isinstance(ahd, (list, tuple)) else (ahd,) * len(boc)
    return UNetConfig(in_channels=raw.get("in_channels", 4), out_channels=raw.get("out_channels", 4),
                      block_out_channels=boc, layers_per_block=raw.get("layers_per_block", 2),
                      cross_attention_dim=raw.get("cross_attention_dim", 768), attention_head_dim=ahd,
                      down_has_attn=has_attn, norm_num_groups=raw.get("norm_num_groups", 32),
                      norm_eps=raw.get("norm_eps", 1e-5),
                      use_linear_projection=raw.get("use_linear_projection", False),
                      sample_size=raw.get("sample_size", 64))


def config_to_dict(cfg: UNetConfig) -> dict:
    return {
        "_class_name": "UNet2DConditionModel", "act_fn": "silu", "in_channels": cfg.in_channels,
        "out_channels": cfg.out_channels, "block_out_channels": list(cfg.block_out_channels),
        "layers_per_block": cfg.layers_per_block, "cross_attention_dim": cfg.cross_attention_dim,
        "attention_head_dim": list(cfg.attention_head_dim),
        "down_block_types": ["CrossAttnDownBlock2D" if a else "DownBlock2D" for a in cfg.down_has_attn],
        "up_block_types": ["CrossAttnUpBlock2D" if a else "UpBlock2D" for a in reversed(cfg.down_has_attn)],
        "mid_block_type": "UNetMidBlock2DCrossAttn", "norm_num_groups": cfg.norm_num_groups,
        "norm_eps": cfg.norm_eps, "use_linear_projection": cfg.use_linear_projection,
        "sample_size": cfg.sample_size, "flip_sin_to_cos": True, "freq_shift": 0, "downsample_padding": 1,
        "center_input_sample": False,
    }


class _UNetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ehs, engine, sample, timesteps):
        ctx.engine = engine
        ctx.ehs_dtype = ehs.dtype
        return engine.forward(sample, timesteps, ehs.to(POLICY.act), save_for_backward=True)

    @staticmethod
    def backward(ctx, d_out):
        d_ehs = ctx.engine.backward(d_out.to(POLICY.act).contiguous())
        return d_ehs.to(ctx.ehs_dtype), None, None, None


class UNet2DConditionModel:
    def __init__(self, config: UNetConfig, state_dict: Dict[str, torch.Tensor], raw_config: Optional[dict] = None):
        self._cfg = config
        self._raw = dict(raw_config) if raw_config is not None else config_to_dict(config)
        self.config = SimpleNamespace(**self._raw)
        self._sd = {k: v.detach() for k, v in state_dict.items()}
        self._engine: Optional[UNetEngine] = None
        self._device = torch.device("cpu")
        self._dtype = next(iter(self._sd.values())).dtype if self._sd else F32
        self.training = True

    # ------------------------------------------------------------------ load / save
    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder: Optional[str] = None, revision=None,
                        variant: Optional[str] = None, **_unused):
        d = os.path.join(pretrained_model_name_or_path, subfolder) if subfolder else pretrained_model_name_or_path
        with open(os.path.join(d, "config.json")) as f:
            raw = json.load(f)
        stem = "diffusion_pytorch_model" + (f".{variant}" if variant else "")
        st = os.path.join(d, stem + ".safetensors")
        if os.path.exists(st):
            from safetensors.torch import load_file
            sd = load_file(st)
        elif os.path.exists(os.path.join(d, stem + ".bin")):
            sd = torch.load(os.path.join(d, stem + ".bin"), map_location="cpu", weights_only=True)
        else:
            raise OSError(f"no {stem}.safetensors / .bin under {d}")
        return cls(_engine_config(raw), sd, raw)

    def save_pretrained(self, save_directory: str, safe_serialization: bool = True, variant: Optional[str] = None,
                        **_unused):
        os.makedirs(save_directory, exist_ok=True)
        sd = {k: v.detach().cpu().contiguous() for k, v in self._sd.items()}
        stem = "diffusion_pytorch_model" + (f".{variant}" if variant else "")
        if safe_serialization:
            from safetensors.torch import save_file
            save_file(sd, os.path.join(save_directory, stem + ".safetensors"), metadata={"format": "pt"})
        else:
            torch.save(sd, os.path.join(save_directory, stem + ".bin"))
        with open(os.path.join(save_directory, "config.json"), "w") as f:
            json.dump(self._raw, f, indent=2, sort_keys=True)

    # ------------------------------------------------------------------ nn.Module-like plumbing
    def eval(self):
        self.training = False
        return self

    def train(self, mode: bool = True):
        if mode:
            raise NotImplementedError("the UNet is frozen on the TextBoost path (--unet_params_to_train none)")
        return self.eval()

    def requires_grad_(self, flag: bool = False):
        if flag:
            raise NotImplementedError("UNet weight gradients are not part of this path (dgrad only)")
        return self

    def state_dict(self):
        return dict(self._sd)

    def named_parameters(self):
        yield from self._sd.items()

    def parameters(self):
        yield from self._sd.values()

    @property
    def device(self):
        return self._device

    @property
    def dtype(self):
        return self._dtype

    def to(self, device=None, dtype=None):
        if isinstance(device, torch.dtype):
            device, dtype = None, device
        if dtype is not None:
            if dtype not in (POLICY.act, F32):
                raise NotImplementedError(f"the B200 UNet computes in {POLICY.act} (precision policy {POLICY.name}); "
                                          f"asked for {dtype}")
            self._dtype = dtype
        if device is not None:
            device = torch.device(device)
            if device.type == "cuda":
                if self._engine is None:
                    dsd = {k: v.to(device=device, dtype=POLICY.act) for k, v in self._sd.items()}
                    self._engine = UNetEngine(self._cfg, dsd)
                    del dsd
            elif self._engine is not None:
                raise RuntimeError("the B200 UNet cannot be moved off the GPU (no CPU path)")
            self._device = device
        return self

    def cuda(self, device=None):
        return self.to(torch.device("cuda", torch.cuda.current_device() if device is None else device))

    @property
    def engine(self) -> UNetEngine:
        if self._engine is None:
            raise RuntimeError("UNet2DConditionModel has no CPU execution path: move it to an sm_100 device "
                               "first (.to('cuda')); the CUDA extension is the product")
        return self._engine

    # ------------------------------------------------------------------ forward
    def forward(self, sample, timestep, encoder_hidden_states, return_dict: bool = True, **unsupported):
        extra = {k: v for k, v in unsupported.items() if v is not None}
        if extra:
            raise NotImplementedError(f"arguments {sorted(extra)} are not used by the TextBoost path")
        e = self.engine
        B = sample.shape[0]
        dev = sample.device
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([timestep], device=dev)
        t = timestep.to(device=dev, dtype=torch.int64).reshape(-1)
        if t.numel() == 1 and B > 1:
            t = t.expand(B)
        x = sample.to(POLICY.act).contiguous()
        if encoder_hidden_states.requires_grad and torch.is_grad_enabled():
            out = _UNetFunction.apply(encoder_hidden_states, e, x, t.contiguous())
        else:
            out = e.forward(x, t.contiguous(), encoder_hidden_states.to(POLICY.act), save_for_backward=False)
        if self._dtype == F32:
            out = out.float()
        return UNet2DConditionOutput(out) if return_dict else (out,)

    __call__ = forward


__all__ = ["UNet2DConditionModel", "UNet2DConditionOutput", "config_to_dict"]
