"""Frozen AutoencoderKL encoder on the B200 kernels: pixels -> training latents (forward only, no autograd).

Drop-in for the reference's ``vae.encode(pixel_values).latent_dist.sample() * vae.config.scaling_factor``
(/root/reference/train_textboost.py:651-653 load, :697 freeze, :938 device move, :1027 pixel cast, :1036-1037 call);
module graph = diffusers 0.29 ``AutoencoderKL.encoder`` + ``quant_conv`` with the diffusers state-dict keys
(restated for the tests in oracle/vae_ref.py).  SURVEY.md §8 f1, image half.

Design:
  * activations channels-last fp16, every conv / linear one tcgen05 GEMM with fp32 accumulation (the reference keeps
    the VAE in fp32: the parity tolerance for this engine is the fp16 one written in tests/test_gpu_vae.py);
  * conv_in (3 -> 128) is the direct small-channel kernel; the three Downsample2D layers (right/bottom padding,
    stride 2) are a window gather + GEMM; GroupNorm(+SiLU) is one HBM pass; the residual add is a conv epilogue;
  * the mid-block attention has ONE head of 512 channels over H/8*W/8 tokens, outside the flash kernel's head sizes:
    it runs as GEMMs per image (QK^T with the 1/sqrt(C) scale in the epilogue, row softmax in place, P V with V
    produced already transposed by W_v h^T; the V bias is added after P V, exact because softmax rows sum to one);
  * conv_out (512 -> 8) and quant_conv (1x1, 8 -> 8) are folded into one 3x3 weight at load time (fp32 on the host
    side of the load, then cast) and zero-padded to 64 output channels so they run on the tensor-core conv;
  * images are processed in chunks (``max_chunk``) so the 512^2 level-0 activations stay a few hundred MB.
"""
from __future__ import annotations

import dataclasses
import json
import os
from typing import Dict, Optional, Tuple

import torch

from . import ops
from .unet import _Conv3, _Linear, _conv_fwd_weight

from .precision import POLICY
F32 = torch.float32


@dataclasses.dataclass
class VAEConfig:
    in_channels: int = 3
    latent_channels: int = 4
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    scaling_factor: float = 0.18215

    @staticmethod
    def from_dict(d) -> "VAEConfig":
        return VAEConfig(in_channels=d.get("in_channels", 3), latent_channels=d.get("latent_channels", 4),
                         block_out_channels=tuple(d.get("block_out_channels", (128, 256, 512, 512))),
                         layers_per_block=d.get("layers_per_block", 2),
                         norm_num_groups=d.get("norm_num_groups", 32),
                         scaling_factor=d.get("scaling_factor", 0.18215))


_EPS = 1e-6


def config_to_dict(cfg: VAEConfig) -> dict:
    """The keys of a diffusers ``vae/config.json`` this engine reads."""
    return {"_class_name": "AutoencoderKL", "in_channels": cfg.in_channels, "latent_channels": cfg.latent_channels,
            "block_out_channels": list(cfg.block_out_channels), "layers_per_block": cfg.layers_per_block,
            "norm_num_groups": cfg.norm_num_groups, "scaling_factor": cfg.scaling_factor}


def vae_encoder_shapes(cfg: VAEConfig) -> Dict[str, Tuple[int, ...]]:
    """diffusers state-dict keys and shapes of ``AutoencoderKL.encoder`` + ``quant_conv`` (what the engine loads)."""
    ch = cfg.block_out_channels
    L2 = 2 * cfg.latent_channels
    out: Dict[str, Tuple[int, ...]] = {}

    def conv(name, cout, cin, k):
        out[name + ".weight"], out[name + ".bias"] = (cout, cin, k, k), (cout,)

    def vec(name, c):
        out[name + ".weight"], out[name + ".bias"] = (c,), (c,)

    def resnet(p, cin, cout):
        vec(p + "norm1", cin)
        conv(p + "conv1", cout, cin, 3)
        vec(p + "norm2", cout)
        conv(p + "conv2", cout, cout, 3)
        if cin != cout:
            conv(p + "conv_shortcut", cout, cin, 1)

    conv("encoder.conv_in", ch[0], cfg.in_channels, 3)
    cin = ch[0]
    for i, c in enumerate(ch):
        for j in range(cfg.layers_per_block):
            resnet(f"encoder.down_blocks.{i}.resnets.{j}.", cin if j == 0 else c, c)
        if i != len(ch) - 1:
            conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", c, c, 3)
        cin = c
    c = ch[-1]
    resnet("encoder.mid_block.resnets.0.", c, c)
    resnet("encoder.mid_block.resnets.1.", c, c)
    a = "encoder.mid_block.attentions.0."
    vec(a + "group_norm", c)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        out[a + n + ".weight"], out[a + n + ".bias"] = (c, c), (c,)
    vec("encoder.conv_norm_out", c)
    conv("encoder.conv_out", L2, c, 3)
    conv("quant_conv", L2, L2, 1)
    return out


def vae_decoder_shapes(cfg: VAEConfig) -> Dict[str, Tuple[int, ...]]:
    """diffusers state-dict keys and shapes of ``AutoencoderKL.decoder`` + ``post_quant_conv``."""
    ch = tuple(reversed(cfg.block_out_channels))
    L = cfg.latent_channels
    out: Dict[str, Tuple[int, ...]] = {}

    def conv(name, cout, cin, k):
        out[name + ".weight"], out[name + ".bias"] = (cout, cin, k, k), (cout,)

    def vec(name, c):
        out[name + ".weight"], out[name + ".bias"] = (c,), (c,)

    def resnet(p, cin, cout):
        vec(p + "norm1", cin)
        conv(p + "conv1", cout, cin, 3)
        vec(p + "norm2", cout)
        conv(p + "conv2", cout, cout, 3)
        if cin != cout:
            conv(p + "conv_shortcut", cout, cin, 1)

    conv("decoder.conv_in", ch[0], L, 3)
    c = ch[0]
    resnet("decoder.mid_block.resnets.0.", c, c)
    resnet("decoder.mid_block.resnets.1.", c, c)
    a = "decoder.mid_block.attentions.0."
    vec(a + "group_norm", c)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        out[a + n + ".weight"], out[a + n + ".bias"] = (c, c), (c,)
    cin = ch[0]
    for i, c in enumerate(ch):
        for j in range(cfg.layers_per_block + 1):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}.", cin if j == 0 else c, c)
        if i != len(ch) - 1:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", c, c, 3)
        cin = c
    vec("decoder.conv_norm_out", ch[-1])
    conv("decoder.conv_out", cfg.in_channels, ch[-1], 3)
    conv("post_quant_conv", L, L, 1)
    return out


class _Resnet:
    def __init__(self, sd, p, groups):
        self.G = groups
        self.n1 = (sd[p + "norm1.weight"], sd[p + "norm1.bias"])
        self.c1 = _Conv3(sd[p + "conv1.weight"], sd[p + "conv1.bias"])
        self.n2 = (sd[p + "norm2.weight"], sd[p + "norm2.bias"])
        self.c2 = _Conv3(sd[p + "conv2.weight"], sd[p + "conv2.bias"])
        self.c1.wd = self.c2.wd = None  # forward only: drop the dgrad copies
        self.sc = None
        if p + "conv_shortcut.weight" in sd:
            self.sc = _Linear(sd[p + "conv_shortcut.weight"], sd[p + "conv_shortcut.bias"])
            self.sc.wt = None

    def forward(self, x):
        B, H, W, Cin = x.shape
        h, _ = ops.groupnorm(x, *self.n1, self.G, _EPS, True)
        h = self.c1.fwd(h)
        h, _ = ops.groupnorm(h, *self.n2, self.G, _EPS, True)
        res = x if self.sc is None else self.sc.fwd(x.view(-1, Cin)).view(B, H, W, -1)
        return self.c2.fwd(h, residual=res)


class _Down:
    def __init__(self, sd, p):
        self.wk = _conv_fwd_weight(sd[p + "conv.weight"])
        self.b = sd[p + "conv.bias"]

    def forward(self, x):
        B, H, W, _ = x.shape
        col = ops.im2col3x3s2_pad(x, 0)
        return ops.gemm(col, self.wk, bias=self.b).view(B, H // 2, W // 2, -1)


class _MidAttention:
    def __init__(self, sd, p, groups):
        self.G = groups
        self.n = (sd[p + "group_norm.weight"], sd[p + "group_norm.bias"])
        self.wqk = torch.cat([sd[p + "to_q.weight"], sd[p + "to_k.weight"]], 0).contiguous()
        self.bqk = torch.cat([sd[p + "to_q.bias"], sd[p + "to_k.bias"]], 0).contiguous()
        self.wv, self.bv = sd[p + "to_v.weight"].contiguous(), sd[p + "to_v.bias"]
        self.wo, self.bo = sd[p + "to_out.0.weight"].contiguous(), sd[p + "to_out.0.bias"]

    def forward(self, x):
        B, H, W, Cc = x.shape
        N = H * W
        h, _ = ops.groupnorm(x, *self.n, self.G, _EPS, False)
        h = h.view(B, N, Cc)
        qk = ops.gemm(h.view(B * N, Cc), self.wqk, bias=self.bqk).view(B, N, 2 * Cc)
        o = torch.empty((B, N, Cc), device=x.device, dtype=POLICY.act)
        for b in range(B):
            s = ops.gemm(qk[b, :, :Cc], qk[b, :, Cc:], alpha=Cc ** -0.5)      # [N, N] scores
            ops.softmax_rows_(s)
            vt = ops.gemm(self.wv, h[b])                                       # V^T [C, N] (bias deferred)
            ops.gemm(s, vt, bias=self.bv, out=o[b])
        return ops.gemm(o.view(B * N, Cc), self.wo, bias=self.bo, residual=x.view(B * N, Cc)).view(B, H, W, Cc)


class VAEEncoderEngine:
    """Weights + forward orchestration.  `sd` maps diffusers AutoencoderKL keys (encoder.*, quant_conv.*) to tensors
    on the CUDA device; decoder / post_quant_conv keys are ignored."""

    PAD_OUT = 64  # conv_out channels padded to the narrowest tensor-core tile

    def __init__(self, cfg: VAEConfig, sd: Dict[str, torch.Tensor], max_chunk: int = 4):
        self.cfg = cfg
        self.max_chunk = max_chunk
        G = cfg.norm_num_groups
        ch = cfg.block_out_channels
        L2 = 2 * cfg.latent_channels
        # conv_out followed by the 1x1 quant_conv is one 3x3 conv: W' = Wq . Wc, b' = Wq . bc + bq (fp32, then fp16)
        wq = sd["quant_conv.weight"].detach().to(F32).reshape(L2, L2)
        wc = sd["encoder.conv_out.weight"].detach().to(F32)
        w_fold = torch.einsum("om,mikl->oikl", wq, wc)
        b_fold = wq @ sd["encoder.conv_out.bias"].detach().to(F32) + sd["quant_conv.bias"].detach().to(F32)
        dev = wc.device
        w_pad = torch.zeros((self.PAD_OUT,) + tuple(wc.shape[1:]), device=dev, dtype=F32)
        w_pad[:L2] = w_fold
        b_pad = torch.zeros(self.PAD_OUT, device=dev, dtype=F32)
        b_pad[:L2] = b_fold
        self.out_w = _conv_fwd_weight(w_pad.to(POLICY.act))
        self.out_b = b_pad.to(POLICY.act)
        sd = {k[len("encoder."):]: v.detach().to(dtype=POLICY.act).contiguous() for k, v in sd.items()
              if k.startswith("encoder.")}
        self.conv_in_w, self.conv_in_b = sd["conv_in.weight"], sd["conv_in.bias"]
        self.down = []
        for i in range(len(ch)):
            p = f"down_blocks.{i}."
            res = [_Resnet(sd, f"{p}resnets.{j}.", G) for j in range(cfg.layers_per_block)]
            down = _Down(sd, f"{p}downsamplers.0.") if i != len(ch) - 1 else None
            self.down.append((res, down))
        self.mid = (_Resnet(sd, "mid_block.resnets.0.", G), _MidAttention(sd, "mid_block.attentions.0.", G),
                    _Resnet(sd, "mid_block.resnets.1.", G))
        self.norm_out = (sd["conv_norm_out.weight"], sd["conv_norm_out.bias"])

    @property
    def downscale(self) -> int:
        return 2 ** (len(self.cfg.block_out_channels) - 1)

    def _moments_rows(self, pixels):
        """pixels [b,3,H,W] fp16 NCHW -> fp16 [b*h*w, PAD_OUT] rows whose first 2L columns are mean | logvar."""
        x = ops.conv_in(pixels, self.conv_in_w, self.conv_in_b)
        for res, down in self.down:
            for r in res:
                x = r.forward(x)
            if down is not None:
                x = down.forward(x)
        r0, attn, r1 = self.mid
        x = r1.forward(attn.forward(r0.forward(x)))
        h, _ = ops.groupnorm(x, *self.norm_out, self.cfg.norm_num_groups, _EPS, True)
        return ops.conv3x3(h, self.out_w, bias=self.out_b).view(-1, self.PAD_OUT)

    def _check(self, pixels):
        if pixels.dim() != 4 or pixels.shape[1] != self.cfg.in_channels:
            raise ValueError(f"expected pixel_values [B,{self.cfg.in_channels},H,W], got {tuple(pixels.shape)}")
        f = self.downscale
        if pixels.shape[2] % f or pixels.shape[3] % f:
            raise ValueError(f"image size {tuple(pixels.shape[2:])} must be a multiple of {f}")
        if not pixels.is_cuda:
            raise RuntimeError("VAEEncoderEngine runs on the CUDA device only (no CPU path)")

    def moments(self, pixel_values):
        """-> (mean, std) fp32 [B, L, H/8, W/8] of the diagonal Gaussian posterior."""
        self._check(pixel_values)
        B, _, H, W = pixel_values.shape
        f, L = self.downscale, self.cfg.latent_channels
        hw = (H // f) * (W // f)
        means, stds = [], []
        for i in range(0, B, self.max_chunk):
            px = pixel_values[i:i + self.max_chunk].to(POLICY.act).contiguous()
            rows = self._moments_rows(px)
            _, m, s = ops.vae_sample(rows, px.shape[0], hw, L, want_moments=True)
            means.append(m)
            stds.append(s)
        return (torch.cat(means).view(B, L, H // f, W // f), torch.cat(stds).view(B, L, H // f, W // f))

    def encode_latents(self, pixel_values, eps: Optional[torch.Tensor] = None, generator=None):
        """(mean + std * eps) * scaling_factor, fp32 [B, L, H/8, W/8]; eps defaults to torch.randn on the device
        (what DiagonalGaussianDistribution.sample draws)."""
        self._check(pixel_values)
        B, _, H, W = pixel_values.shape
        f, L = self.downscale, self.cfg.latent_channels
        h, w = H // f, W // f
        if eps is None:
            eps = torch.randn((B, L, h, w), device=pixel_values.device, dtype=F32, generator=generator)
        eps = eps.to(F32).contiguous()
        out = []
        for i in range(0, B, self.max_chunk):
            px = pixel_values[i:i + self.max_chunk].to(POLICY.act).contiguous()
            rows = self._moments_rows(px)
            lat, _, _ = ops.vae_sample(rows, px.shape[0], h * w, L, eps=eps[i:i + self.max_chunk],
                                       scaling_factor=self.cfg.scaling_factor)
            out.append(lat)
        return torch.cat(out).view(B, L, h, w)


class _Up:
    """diffusers Upsample2D: nearest-neighbour 2x, then conv3x3."""

    def __init__(self, sd, p):
        self.wk = _conv_fwd_weight(sd[p + "conv.weight"])
        self.b = sd[p + "conv.bias"]

    def forward(self, x):
        return ops.conv3x3(ops.upsample2x(x), self.wk, bias=self.b)


class VAEDecoderEngine:
    """``vae.decode(latents / scaling_factor).sample`` + the image post-processing of StableDiffusionPipeline, forward
    only: the last stage of the validation / inference sampler (train_textboost.py:512-513,
    /root/reference/inference.py:99-105).  `sd` maps diffusers keys (decoder.*, post_quant_conv.*) to CUDA tensors.

    post_quant_conv (1x1 over 4 channels) runs fused with the 1/scaling_factor and the fp16 cast in front of conv_in
    (it cannot be folded into conv_in's weights: its bias would leak into the zero padding); conv_out (128 -> 3) is
    zero-padded to 64 output channels for the tensor-core conv and only the first three columns are read back."""

    PAD_OUT = 64

    def __init__(self, cfg: VAEConfig, sd: Dict[str, torch.Tensor], max_chunk: int = 4):
        self.cfg, self.max_chunk = cfg, max_chunk
        G = cfg.norm_num_groups
        ch = tuple(reversed(cfg.block_out_channels))
        L = cfg.latent_channels
        self.pq_w = sd["post_quant_conv.weight"].detach().to(F32).reshape(L, L).contiguous()
        self.pq_b = sd["post_quant_conv.bias"].detach().to(F32).contiguous()
        sd = {k[len("decoder."):]: v.detach().to(dtype=POLICY.act).contiguous() for k, v in sd.items()
              if k.startswith("decoder.")}
        self.conv_in_w, self.conv_in_b = sd["conv_in.weight"], sd["conv_in.bias"]
        self.mid = (_Resnet(sd, "mid_block.resnets.0.", G), _MidAttention(sd, "mid_block.attentions.0.", G),
                    _Resnet(sd, "mid_block.resnets.1.", G))
        self.up = []
        for i in range(len(ch)):
            p = f"up_blocks.{i}."
            res = [_Resnet(sd, f"{p}resnets.{j}.", G) for j in range(cfg.layers_per_block + 1)]
            up = _Up(sd, f"{p}upsamplers.0.") if i != len(ch) - 1 else None
            self.up.append((res, up))
        self.norm_out = (sd["conv_norm_out.weight"], sd["conv_norm_out.bias"])
        w = sd["conv_out.weight"]
        w_pad = torch.zeros((self.PAD_OUT,) + tuple(w.shape[1:]), device=w.device, dtype=POLICY.act)
        w_pad[:cfg.in_channels] = w
        b_pad = torch.zeros(self.PAD_OUT, device=w.device, dtype=POLICY.act)
        b_pad[:cfg.in_channels] = sd["conv_out.bias"]
        self.out_w, self.out_b = _conv_fwd_weight(w_pad), b_pad

    @property
    def upscale(self) -> int:
        return 2 ** (len(self.cfg.block_out_channels) - 1)

    def _image_rows(self, latents, scaling_factor):
        """latents fp32 [b,L,h,w] -> fp16 rows [b*H*W, PAD_OUT]; columns [0,3) are the image in [-1,1] (unclamped)."""
        z = ops.vae_decode_in(latents, self.pq_w, self.pq_b, scaling_factor)
        x = ops.conv_in(z, self.conv_in_w, self.conv_in_b)
        r0, attn, r1 = self.mid
        x = r1.forward(attn.forward(r0.forward(x)))
        for res, up in self.up:
            for r in res:
                x = r.forward(x)
            if up is not None:
                x = up.forward(x)
        h, _ = ops.groupnorm(x, *self.norm_out, self.cfg.norm_num_groups, _EPS, True)
        return ops.conv3x3(h, self.out_w, bias=self.out_b).view(-1, self.PAD_OUT)

    def _check(self, latents):
        if latents.dim() != 4 or latents.shape[1] != self.cfg.latent_channels:
            raise ValueError(f"expected latents [B,{self.cfg.latent_channels},h,w], got {tuple(latents.shape)}")
        if not latents.is_cuda:
            raise RuntimeError("VAEDecoderEngine runs on the CUDA device only (no CPU path)")

    def _chunks(self, latents):
        self._check(latents)
        latents = latents.to(F32).contiguous()
        for i in range(0, latents.shape[0], self.max_chunk):
            yield latents[i:i + self.max_chunk]

    def decode(self, z, scaling_factor: float = 1.0):
        """diffusers ``vae.decode(z).sample`` for z = latents / scaling_factor (or pass latents and the factor):
        fp16 [B,3,H,W], unclamped."""
        f, Cimg = self.upscale, self.cfg.in_channels
        out = []
        for lat in self._chunks(z):
            b, _, h, w = lat.shape
            rows = self._image_rows(lat, scaling_factor)
            out.append(rows[:, :Cimg].reshape(b, h * f, w * f, Cimg).permute(0, 3, 1, 2))
        return torch.cat(out)

    def decode_u8(self, latents):
        """Scaled latents (as the sampler leaves them) -> uint8 [B,H,W,3] images: decode(latents / scaling_factor),
        (x/2 + .5).clamp(0,1) * 255 rounded."""
        f, Cimg = self.upscale, self.cfg.in_channels
        out = []
        for lat in self._chunks(latents):
            b, _, h, w = lat.shape
            rows = self._image_rows(lat, self.cfg.scaling_factor)
            out.append(ops.image_u8(rows, rows.shape[0], Cimg).view(b, h * f, w * f, Cimg))
        return torch.cat(out)


# --------------------------------------------------------------------------------------------------------------
# Host mirror of the slice of diffusers.AutoencoderKL the reference touches.
class _Posterior:
    """``latent_dist``: sample() / mode() / mean / std, as DiagonalGaussianDistribution exposes them."""

    def __init__(self, engine: VAEEncoderEngine, pixel_values):
        self._engine, self._px = engine, pixel_values
        self._moments = None

    def _m(self):
        if self._moments is None:
            self._moments = self._engine.moments(self._px)
        return self._moments

    @property
    def mean(self):
        return self._m()[0]

    @property
    def std(self):
        return self._m()[1]

    def mode(self):
        return self.mean

    def sample(self, generator=None):
        # latents / scaling_factor: the caller multiplies by config.scaling_factor itself (train_textboost.py:1037)
        mean, std = self._m()
        eps = torch.randn(mean.shape, device=mean.device, dtype=F32, generator=generator)
        return mean + std * eps


class _EncoderOutput:
    def __init__(self, latent_dist):
        self.latent_dist = latent_dist


class _DecoderOutput:
    def __init__(self, sample):
        self.sample = sample


class _Config(dict):
    __getattr__ = dict.__getitem__


class AutoencoderKL:
    """``AutoencoderKL.from_pretrained(path, subfolder="vae")`` / ``.encode(x).latent_dist.sample()`` /
    ``.config.scaling_factor`` / ``.dtype`` / ``.to(device, dtype=...)`` / ``.eval()`` / ``.requires_grad_(False)``."""

    def __init__(self, cfg: VAEConfig, state_dict: Dict[str, torch.Tensor]):
        self._cfg = cfg
        self.config = _Config(dataclasses.asdict(cfg))
        self._sd = dict(state_dict)
        self.engine: Optional[VAEEncoderEngine] = None
        self.decoder_engine: Optional[VAEDecoderEngine] = None
        self.dtype = torch.float32  # what callers cast pixel_values to (train_textboost.py:1027)

    @classmethod
    def from_pretrained(cls, path, subfolder=None, revision=None, variant=None, **_):
        root = os.path.join(path, subfolder) if subfolder else path
        with open(os.path.join(root, "config.json")) as f:
            cfg = VAEConfig.from_dict(json.load(f))
        stem = "diffusion_pytorch_model" + (f".{variant}" if variant else "")
        st = os.path.join(root, stem + ".safetensors")
        if os.path.exists(st):
            from safetensors.torch import load_file
            sd = load_file(st)
        elif os.path.exists(os.path.join(root, stem + ".bin")):
            sd = torch.load(os.path.join(root, stem + ".bin"), map_location="cpu", weights_only=True)
        else:
            raise OSError(f"no {stem}.safetensors / .bin under {root}")
        return cls(cfg, sd)

    def eval(self):
        return self

    def requires_grad_(self, flag=False):
        if flag:
            raise NotImplementedError("the VAE is frozen on this path (train_textboost.py:697)")
        return self

    def to(self, device=None, dtype=None):
        """dtype is accepted for call-site compatibility (train_textboost.py:938 asks for fp32) but the engines compute
        in fp16 with fp32 accumulation whatever it says; asking for anything else is reported once, with the measured
        deviation, rather than dropped silently."""
        if dtype is not None and dtype != POLICY.act and not getattr(AutoencoderKL, "_dtype_warned", False):
            import warnings
            AutoencoderKL._dtype_warned = True
            warnings.warn(f"AutoencoderKL.to(dtype={dtype}): the B200 VAE engines compute in fp16 with fp32 accumulation "
                          "(measured against the fp32 oracle at 512^2: 1.6e-3 relative L2 on the posterior mean, 4.7e-4 "
                          "on the scaled latents; DESIGN.md section 7); the requested dtype is not applied")
        if device is not None and torch.device(device).type == "cuda":
            sd = {k: v.to(device) for k, v in self._sd.items()}
            if "encoder.conv_in.weight" in sd:
                self.engine = VAEEncoderEngine(self._cfg, sd)
            if "decoder.conv_in.weight" in sd:
                self.decoder_engine = VAEDecoderEngine(self._cfg, sd)
        return self

    def encode(self, pixel_values):
        if self.engine is None:
            raise RuntimeError("AutoencoderKL.encode: call .to('cuda') first (no CPU path), on a checkpoint that has "
                               "the encoder weights")
        return _EncoderOutput(_Posterior(self.engine, pixel_values))

    def decode(self, z, return_dict=True):
        """``vae.decode(latents / scaling_factor)`` -> object with ``.sample`` fp16 [B,3,H,W] (or a 1-tuple)."""
        if self.decoder_engine is None:
            raise RuntimeError("AutoencoderKL.decode: call .to('cuda') first (no CPU path), on a checkpoint that has "
                               "the decoder weights")
        sample = self.decoder_engine.decode(z)
        return _DecoderOutput(sample) if return_dict else (sample,)
