"""oracle (test infrastructure): diffusers 0.29.0 ``UNet2DConditionModel`` restated in plain PyTorch.

diffusers is NOT installed and cannot be (no network): this file restates the published
architecture of the modules the reference instantiates at
/root/reference/train_textboost.py:654-656 and calls at :1063-1067 —
``models/unets/unet_2d_condition.py``, ``unet_2d_blocks.py`` (CrossAttnDownBlock2D, DownBlock2D,
UNetMidBlock2DCrossAttn, UpBlock2D, CrossAttnUpBlock2D), ``models/resnet.py`` (ResnetBlock2D),
``downsampling.py`` / ``upsampling.py``, ``transformers/transformer_2d.py``, ``attention.py``
(BasicTransformerBlock, FeedForward, GEGLU), ``attention_processor.py`` (AttnProcessor2_0) and
``embeddings.py`` (get_timestep_embedding, TimestepEmbedding).  Module names reproduce the
diffusers state-dict keys (Appendix A.4), so a real ``diffusion_pytorch_model.safetensors`` loads
with ``load_state_dict``.  **Parity of this file against diffusers itself is unpinned** — there is
no executable diffusers here; the parameter count (859,520,964 for SD-1.5) is the one check.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Tuple

import torch
import torch.nn.functional as F
from torch import nn


@dataclasses.dataclass
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    cross_attention_dim: int = 768
    attention_head_dim: Tuple[int, ...] = (8, 8, 8, 8)  # = number of heads per level (diffusers quirk)
    down_has_attn: Tuple[bool, ...] = (True, True, True, False)
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    use_linear_projection: bool = False
    sample_size: int = 64

    @staticmethod
    def sd15() -> "UNetConfig":
        return UNetConfig()

    @staticmethod
    def sd21() -> "UNetConfig":
        return UNetConfig(cross_attention_dim=1024, attention_head_dim=(5, 10, 20, 20),
                          use_linear_projection=True, sample_size=96)

    @staticmethod
    def tiny(cross_attention_dim: int = 64) -> "UNetConfig":
        """Same topology, small widths: used by CPU-speed parity tests."""
        return UNetConfig(block_out_channels=(64, 128, 128, 128), attention_head_dim=(2, 2, 2, 2),
                          cross_attention_dim=cross_attention_dim, sample_size=16)


def timestep_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    """embeddings.get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0), fp32."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None, :]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


class TimestepEmbedding(nn.Module):
    def __init__(self, cin, dim):
        super().__init__()
        self.linear_1 = nn.Linear(cin, dim)
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, temb_ch, groups, eps):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_ch, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Attention(nn.Module):
    def __init__(self, dim, heads, ctx_dim=None):
        super().__init__()
        self.heads = heads
        ctx_dim = dim if ctx_dim is None else ctx_dim
        self.to_q = nn.Linear(dim, dim, bias=False)
        self.to_k = nn.Linear(ctx_dim, dim, bias=False)
        self.to_v = nn.Linear(ctx_dim, dim, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Identity()])

    def forward(self, x, ctx=None):
        ctx = x if ctx is None else ctx
        B, N, C = x.shape
        d = C // self.heads
        q = self.to_q(x).view(B, N, self.heads, d).transpose(1, 2)
        k = self.to_k(ctx).view(B, -1, self.heads, d).transpose(1, 2)
        v = self.to_v(ctx).view(B, -1, self.heads, d).transpose(1, 2)
        w = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) * d ** -0.5, dim=-1)
        o = torch.matmul(w, v).transpose(1, 2).reshape(B, N, C)
        return self.to_out[0](o)


class GEGLU(nn.Module):
    def __init__(self, dim, inner):
        super().__init__()
        self.proj = nn.Linear(dim, inner * 2)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Identity(), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        return self.net[2](self.net[0](x))


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, ctx_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, heads)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, heads, ctx_dim)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, ctx):
        x = x + self.attn1(self.norm1(x))
        x = x + self.attn2(self.norm2(x), ctx)
        x = x + self.ff(self.norm3(x))
        return x


class Transformer2DModel(nn.Module):
    def __init__(self, dim, heads, ctx_dim, groups, linear_proj):
        super().__init__()
        self.linear_proj = linear_proj
        self.norm = nn.GroupNorm(groups, dim, eps=1e-6)
        if linear_proj:
            self.proj_in = nn.Linear(dim, dim)
            self.proj_out = nn.Linear(dim, dim)
        else:
            self.proj_in = nn.Conv2d(dim, dim, 1)
            self.proj_out = nn.Conv2d(dim, dim, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(dim, heads, ctx_dim)])

    def forward(self, x, ctx):
        B, C, H, W = x.shape
        res = x
        h = self.norm(x)
        if self.linear_proj:
            h = self.proj_in(h.permute(0, 2, 3, 1).reshape(B, H * W, C))
        else:
            h = self.proj_in(h).permute(0, 2, 3, 1).reshape(B, H * W, C)
        for blk in self.transformer_blocks:
            h = blk(h, ctx)
        if self.linear_proj:
            h = self.proj_out(h).reshape(B, H, W, C).permute(0, 3, 1, 2)
        else:
            h = self.proj_out(h.reshape(B, H, W, C).permute(0, 3, 1, 2))
        return h + res


class Downsample2D(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    """CrossAttnDownBlock2D (has_attn) / DownBlock2D."""

    def __init__(self, cfg: UNetConfig, cin, cout, heads, has_attn, add_down, temb_ch):
        super().__init__()
        self.resnets = nn.ModuleList()
        if has_attn:
            self.attentions = nn.ModuleList()
        self.has_attn = has_attn
        for i in range(cfg.layers_per_block):
            self.resnets.append(ResnetBlock2D(cin if i == 0 else cout, cout, temb_ch,
                                              cfg.norm_num_groups, cfg.norm_eps))
            if has_attn:
                self.attentions.append(Transformer2DModel(cout, heads, cfg.cross_attention_dim,
                                                          cfg.norm_num_groups,
                                                          cfg.use_linear_projection))
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_down else None

    def forward(self, x, temb, ctx):
        outs = []
        for i, r in enumerate(self.resnets):
            x = r(x, temb)
            if self.has_attn:
                x = self.attentions[i](x, ctx)
            outs.append(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs.append(x)
        return x, outs


class MidBlock(nn.Module):
    """UNetMidBlock2DCrossAttn."""

    def __init__(self, cfg: UNetConfig, ch, heads, temb_ch):
        super().__init__()
        self.resnets = nn.ModuleList([
            ResnetBlock2D(ch, ch, temb_ch, cfg.norm_num_groups, cfg.norm_eps) for _ in range(2)])
        self.attentions = nn.ModuleList([
            Transformer2DModel(ch, heads, cfg.cross_attention_dim, cfg.norm_num_groups,
                               cfg.use_linear_projection)])

    def forward(self, x, temb, ctx):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x, ctx)
        return self.resnets[1](x, temb)


class UpBlock(nn.Module):
    """CrossAttnUpBlock2D (has_attn) / UpBlock2D."""

    def __init__(self, cfg: UNetConfig, cin, cout, prev_out, heads, has_attn, add_up, temb_ch):
        super().__init__()
        self.resnets = nn.ModuleList()
        if has_attn:
            self.attentions = nn.ModuleList()
        self.has_attn = has_attn
        n = cfg.layers_per_block + 1
        for i in range(n):
            skip_ch = cin if i == n - 1 else cout
            res_in = prev_out if i == 0 else cout
            self.resnets.append(ResnetBlock2D(res_in + skip_ch, cout, temb_ch, cfg.norm_num_groups,
                                              cfg.norm_eps))
            if has_attn:
                self.attentions.append(Transformer2DModel(cout, heads, cfg.cross_attention_dim,
                                                          cfg.norm_num_groups,
                                                          cfg.use_linear_projection))
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, x, skips, temb, ctx):
        for i, r in enumerate(self.resnets):
            x = torch.cat([x, skips.pop()], dim=1)
            x = r(x, temb)
            if self.has_attn:
                x = self.attentions[i](x, ctx)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class UNet2DConditionModelRef(nn.Module):
    def __init__(self, cfg: UNetConfig):
        super().__init__()
        self.config = cfg
        ch = cfg.block_out_channels
        temb_ch = ch[0] * 4
        self.conv_in = nn.Conv2d(cfg.in_channels, ch[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(ch[0], temb_ch)
        self.down_blocks = nn.ModuleList()
        out = ch[0]
        for i, c in enumerate(ch):
            cin, out = out, c
            self.down_blocks.append(DownBlock(cfg, cin, out, cfg.attention_head_dim[i],
                                              cfg.down_has_attn[i], i != len(ch) - 1, temb_ch))
        self.mid_block = MidBlock(cfg, ch[-1], cfg.attention_head_dim[-1], temb_ch)
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(ch))
        rev_heads = list(reversed(cfg.attention_head_dim))
        rev_attn = list(reversed(cfg.down_has_attn))
        out = rev[0]
        for i, c in enumerate(rev):
            prev_out, out = out, c
            cin = rev[min(i + 1, len(ch) - 1)]
            self.up_blocks.append(UpBlock(cfg, cin, out, prev_out, rev_heads[i], rev_attn[i],
                                          i != len(ch) - 1, temb_ch))
        self.conv_norm_out = nn.GroupNorm(cfg.norm_num_groups, ch[0], eps=cfg.norm_eps)
        self.conv_out = nn.Conv2d(ch[0], cfg.out_channels, 3, padding=1)


    def add_cross_kv_lora(self, r: int, lora_alpha=None):
        """``unet.add_adapter(LoraConfig(r, lora_alpha=r, init_lora_weights="gaussian",
        target_modules=["attn2.to_k", "attn2.to_v"]))`` of /root/reference/train_textboost.py:712-720: every
        cross-attention's K and V projections become peft LoRA linears (clip_ref.LoraLinear restates peft's layer);
        only the lora_A / lora_B weights are trainable afterwards.  Returns the wrapped modules in forward order,
        to_k before to_v within a block."""
        from .clip_ref import LoraLinear
        self.requires_grad_(False)
        wrapped = []
        for name, mod in list(self.named_modules()):
            if name.endswith("attn2"):
                for t in ("to_k", "to_v"):
                    lin = LoraLinear(getattr(mod, t), r, lora_alpha if lora_alpha is not None else r)
                    setattr(mod, t, lin)
                    wrapped.append((f"{name}.{t}", lin))
        return wrapped

    def cross_kv_lora_modules(self):
        """(module path, LoraLinear) of every adapter, in the engine's order: down blocks, mid block, up blocks."""
        from .clip_ref import LoraLinear
        return [(n, m) for n, m in self.named_modules() if isinstance(m, LoraLinear)]

    def forward(self, sample, timesteps, encoder_hidden_states):
        """sample [B,4,H,W], timesteps int64 [B], encoder_hidden_states [B,L,ctx] -> [B,4,H,W]."""
        cfg = self.config
        temb = timestep_embedding(timesteps, cfg.block_out_channels[0]).to(sample.dtype)
        temb = self.time_embedding(temb)
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, temb, encoder_hidden_states)
            skips.extend(outs)
        x = self.mid_block(x, temb, encoder_hidden_states)
        for blk in self.up_blocks:
            x = blk(x, skips, temb, encoder_hidden_states)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


def init_unet_(model: UNet2DConditionModelRef, seed: int = 0):
    """Deterministic random init that keeps activations O(1) through ~60 layers:
    conv / linear weights N(0, 1/fan_in), residual-branch outputs damped, norm affine 1 + N(0,.1)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() == 1:
                if "norm" in name and name.endswith("weight"):
                    p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
                else:
                    p.copy_(0.05 * torch.randn(p.shape, generator=g))
                continue
            fan_in = p[0].numel()
            gain = 1.0
            if any(s in name for s in ("conv2.", "to_out.0", "ff.net.2", "proj_out")):
                gain = 0.5
            p.copy_(torch.randn(p.shape, generator=g) * gain / math.sqrt(fan_in))
    return model
