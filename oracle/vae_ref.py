"""oracle (test infrastructure, parity UNPINNED: diffusers is not installable here): the encoder half of diffusers
0.29 ``AutoencoderKL`` as used at /root/reference/train_textboost.py:651-653, 938, 1036-1037

    latents = vae.encode(pixel_values).latent_dist.sample() * vae.config.scaling_factor      (fp32, no grad)

restated in plain PyTorch with the diffusers state-dict key names (``encoder.*``, ``quant_conv.*``), for the image
half of SURVEY.md §8 f1, and its decoder half (``AutoencoderKLRef``) for the validation / inference sampler of §8 f3.  SD-1.x / 2.x VAE config: in 3, latent 4,
block_out_channels (128, 256, 512, 512), 2 resnets per block, GroupNorm(32, eps 1e-6), SiLU, downsample = pad
(0,1,0,1) + conv3x3 stride 2, mid block = resnet, single-head attention over 512 channels, resnet; conv_out -> 8
channels (mean | logvar), quant_conv 1x1; scaling_factor 0.18215.
"""
from __future__ import annotations

import dataclasses
from typing import Tuple

import torch
import torch.nn.functional as F
from torch import nn


@dataclasses.dataclass
class VAEConfig:
    in_channels: int = 3
    latent_channels: int = 4
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    scaling_factor: float = 0.18215


class Resnet(nn.Module):
    def __init__(self, cin, cout, groups):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-6)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-6)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        return (x if self.conv_shortcut is None else self.conv_shortcut(x)) + h


class Downsample(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1)))  # right / bottom only (diffusers Downsample2D, padding=0)


class DownBlock(nn.Module):
    def __init__(self, cin, cout, n, groups, down):
        super().__init__()
        self.resnets = nn.ModuleList([Resnet(cin if j == 0 else cout, cout, groups) for j in range(n)])
        self.downsamplers = nn.ModuleList([Downsample(cout)]) if down else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
        return x


class MidAttention(nn.Module):
    """diffusers Attention(heads=1, residual_connection=True, norm_num_groups=32, bias=True) over H*W tokens."""

    def __init__(self, c, groups):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, c, eps=1e-6)
        self.to_q, self.to_k, self.to_v = nn.Linear(c, c), nn.Linear(c, c), nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c), nn.Identity()])

    def forward(self, x):
        B, C, H, W = x.shape
        h = self.group_norm(x).view(B, C, H * W).transpose(1, 2)
        q, k, v = self.to_q(h), self.to_k(h), self.to_v(h)
        w = torch.softmax(q @ k.transpose(1, 2) * C ** -0.5, dim=-1)
        o = self.to_out[0](w @ v).transpose(1, 2).reshape(B, C, H, W)
        return x + o


class MidBlock(nn.Module):
    def __init__(self, c, groups):
        super().__init__()
        self.resnets = nn.ModuleList([Resnet(c, c, groups), Resnet(c, c, groups)])
        self.attentions = nn.ModuleList([MidAttention(c, groups)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class Encoder(nn.Module):
    def __init__(self, cfg: VAEConfig):
        super().__init__()
        ch = cfg.block_out_channels
        self.conv_in = nn.Conv2d(cfg.in_channels, ch[0], 3, padding=1)
        blocks, cin = [], ch[0]
        for i, c in enumerate(ch):
            blocks.append(DownBlock(cin, c, cfg.layers_per_block, cfg.norm_num_groups, down=i != len(ch) - 1))
            cin = c
        self.down_blocks = nn.ModuleList(blocks)
        self.mid_block = MidBlock(ch[-1], cfg.norm_num_groups)
        self.conv_norm_out = nn.GroupNorm(cfg.norm_num_groups, ch[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(ch[-1], 2 * cfg.latent_channels, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class UpBlock(nn.Module):
    """diffusers UpDecoderBlock2D: layers_per_block + 1 resnets, then nearest 2x upsample + conv3x3 (Upsample2D)."""

    def __init__(self, cin, cout, n, groups, up):
        super().__init__()
        self.resnets = nn.ModuleList([Resnet(cin if j == 0 else cout, cout, groups) for j in range(n)])
        self.upsamplers = nn.ModuleList([Upsample(cout)]) if up else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class Upsample(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class Decoder(nn.Module):
    def __init__(self, cfg: VAEConfig):
        super().__init__()
        ch = tuple(reversed(cfg.block_out_channels))
        self.conv_in = nn.Conv2d(cfg.latent_channels, ch[0], 3, padding=1)
        self.mid_block = MidBlock(ch[0], cfg.norm_num_groups)
        blocks, cin = [], ch[0]
        for i, c in enumerate(ch):
            blocks.append(UpBlock(cin, c, cfg.layers_per_block + 1, cfg.norm_num_groups, up=i != len(ch) - 1))
            cin = c
        self.up_blocks = nn.ModuleList(blocks)
        self.conv_norm_out = nn.GroupNorm(cfg.norm_num_groups, ch[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(ch[-1], cfg.in_channels, 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class AutoencoderKLEncoderRef(nn.Module):
    def __init__(self, cfg: VAEConfig = VAEConfig()):
        super().__init__()
        self.cfg = cfg
        self.encoder = Encoder(cfg)
        self.quant_conv = nn.Conv2d(2 * cfg.latent_channels, 2 * cfg.latent_channels, 1)

    def moments(self, pixel_values):
        """[B,3,H,W] in [-1,1] -> (mean, std) of the diagonal Gaussian, each [B,4,H/8,W/8]."""
        m = self.quant_conv(self.encoder(pixel_values))
        mean, logvar = m.chunk(2, dim=1)
        return mean, torch.exp(0.5 * logvar.clamp(-30.0, 20.0))

    def encode_latents(self, pixel_values, eps):
        """train_textboost.py:1036-1037 with the Gaussian sample made explicit: (mean + std * eps) * scaling_factor."""
        mean, std = self.moments(pixel_values)
        return (mean + std * eps) * self.cfg.scaling_factor


class AutoencoderKLRef(AutoencoderKLEncoderRef):
    """Encoder + decoder (diffusers AutoencoderKL: 83,653,863 parameters for the SD config).  ``decode_latents`` is what
    StableDiffusionPipeline does at the end of sampling (train_textboost.py:512-513 via ``pipeline(...)``, and
    /root/reference/inference.py:99-105): ``vae.decode(latents / scaling_factor).sample`` -> (x/2 + .5).clamp(0, 1)."""

    def __init__(self, cfg: VAEConfig = VAEConfig()):
        super().__init__(cfg)
        self.decoder = Decoder(cfg)
        self.post_quant_conv = nn.Conv2d(cfg.latent_channels, cfg.latent_channels, 1)

    def decode(self, z):
        return self.decoder(self.post_quant_conv(z))

    def decode_latents(self, latents):
        image = self.decode(latents / self.cfg.scaling_factor)
        return (image / 2 + 0.5).clamp(0, 1)

    @staticmethod
    def to_uint8(image01):
        """VaeImageProcessor.postprocess(output_type="pil") up to the PIL wrap: NCHW [0,1] -> NHWC uint8."""
        return (image01.permute(0, 2, 3, 1).float() * 255).round().to(torch.uint8)
