"""oracle (test infrastructure): CLIP text encoder + LoRA + TextBoostModel override, plain PyTorch.

Restates
  * transformers ``models/clip/modeling_clip.py`` CLIPTextEmbeddings / CLIPAttention / CLIPMLP /
    CLIPEncoderLayer / CLIPTextTransformer (installed copy lines 221-590; the pinned 4.40.2 computes
    the same function) — called from /root/reference/textboost/text_encoder.py:62-69;
  * peft 0.13.2 ``tuners/lora/layer.py::Linear.forward`` with ``init_lora_weights="gaussian"``
    (``reset_lora_parameters``: A ~ N(0, (1/r)^2), B = 0), configured at
    /root/reference/train_textboost.py:702-709;
  * /root/reference/textboost/text_encoder.py:17-87 (null-embedding buffer and the two in-place
    overrides of the output).

Module / parameter names follow the HF + peft state-dict layout so that a
reference checkpoint's keys line up one to one.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Sequence

import torch
import torch.nn.functional as F
from torch import nn


@dataclasses.dataclass
class ClipTextConfig:
    vocab_size: int = 49408
    hidden_size: int = 768
    intermediate_size: int = 3072
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    max_position_embeddings: int = 77
    hidden_act: str = "quick_gelu"
    layer_norm_eps: float = 1e-5

    @staticmethod
    def clip_l() -> "ClipTextConfig":  # SD-1.x text encoder
        return ClipTextConfig()

    @staticmethod
    def openclip_h() -> "ClipTextConfig":  # SD-2.x text encoder (23 layers kept by diffusers)
        return ClipTextConfig(hidden_size=1024, intermediate_size=4096, num_hidden_layers=23,
                              num_attention_heads=16, hidden_act="gelu")


def _act(name: str):
    if name == "quick_gelu":
        return lambda x: x * torch.sigmoid(1.702 * x)
    if name == "gelu":
        return F.gelu
    raise ValueError(name)


class LoraLinear(nn.Module):
    """peft LoRA Linear: y = base(x) + scaling * B(A(x)), scaling = lora_alpha / r, dropout = identity."""

    def __init__(self, base: nn.Linear, r: int, lora_alpha: int):
        super().__init__()
        self.base_layer = base
        self.lora_A = nn.ModuleDict({"default": nn.Linear(base.in_features, r, bias=False)})
        self.lora_B = nn.ModuleDict({"default": nn.Linear(r, base.out_features, bias=False)})
        self.scaling = lora_alpha / r
        nn.init.normal_(self.lora_A["default"].weight, std=1.0 / r)  # init_lora_weights="gaussian"
        nn.init.zeros_(self.lora_B["default"].weight)
        base.weight.requires_grad_(False)
        if base.bias is not None:
            base.bias.requires_grad_(False)

    def forward(self, x):
        return self.base_layer(x) + self.scaling * self.lora_B["default"](self.lora_A["default"](x))


class ClipAttention(nn.Module):
    def __init__(self, cfg: ClipTextConfig):
        super().__init__()
        d = cfg.hidden_size
        self.num_heads = cfg.num_attention_heads
        self.head_dim = d // self.num_heads
        self.k_proj = nn.Linear(d, d)
        self.v_proj = nn.Linear(d, d)
        self.q_proj = nn.Linear(d, d)
        self.out_proj = nn.Linear(d, d)

    def forward(self, x, mask):
        B, L, D = x.shape
        q = self.q_proj(x).view(B, L, self.num_heads, self.head_dim).transpose(1, 2)
        k = self.k_proj(x).view(B, L, self.num_heads, self.head_dim).transpose(1, 2)
        v = self.v_proj(x).view(B, L, self.num_heads, self.head_dim).transpose(1, 2)
        w = torch.matmul(q, k.transpose(-1, -2)) * self.head_dim ** -0.5
        w = w + mask
        w = torch.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)
        o = torch.matmul(w, v).transpose(1, 2).reshape(B, L, D)
        return self.out_proj(o)


class ClipMLP(nn.Module):
    def __init__(self, cfg: ClipTextConfig):
        super().__init__()
        self.fc1 = nn.Linear(cfg.hidden_size, cfg.intermediate_size)
        self.fc2 = nn.Linear(cfg.intermediate_size, cfg.hidden_size)
        self.act = _act(cfg.hidden_act)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class ClipEncoderLayer(nn.Module):
    def __init__(self, cfg: ClipTextConfig):
        super().__init__()
        self.self_attn = ClipAttention(cfg)
        self.layer_norm1 = nn.LayerNorm(cfg.hidden_size, eps=cfg.layer_norm_eps)
        self.mlp = ClipMLP(cfg)
        self.layer_norm2 = nn.LayerNorm(cfg.hidden_size, eps=cfg.layer_norm_eps)

    def forward(self, x, mask):
        x = x + self.self_attn(self.layer_norm1(x), mask)
        x = x + self.mlp(self.layer_norm2(x))
        return x


class ClipEncoder(nn.Module):
    def __init__(self, cfg: ClipTextConfig):
        super().__init__()
        self.layers = nn.ModuleList([ClipEncoderLayer(cfg) for _ in range(cfg.num_hidden_layers)])


class ClipEmbeddings(nn.Module):
    def __init__(self, cfg: ClipTextConfig):
        super().__init__()
        self.token_embedding = nn.Embedding(cfg.vocab_size, cfg.hidden_size)
        self.position_embedding = nn.Embedding(cfg.max_position_embeddings, cfg.hidden_size)


class ClipTextTransformer(nn.Module):
    def __init__(self, cfg: ClipTextConfig):
        super().__init__()
        self.embeddings = ClipEmbeddings(cfg)
        self.encoder = ClipEncoder(cfg)
        self.final_layer_norm = nn.LayerNorm(cfg.hidden_size, eps=cfg.layer_norm_eps)

    def forward(self, input_ids):
        B, L = input_ids.shape
        e = self.embeddings
        x = e.token_embedding(input_ids) + e.position_embedding.weight[:L].unsqueeze(0)
        mask = torch.full((L, L), float("-inf"), dtype=x.dtype, device=x.device).triu(1)
        for layer in self.encoder.layers:
            x = layer(x, mask)
        return self.final_layer_norm(x)


class TextBoostModelRef(nn.Module):
    """textboost/text_encoder.py:17-87 restated on top of the CLIP restatement above.

    forward(input_ids) returns the last hidden state [B, L, D] (the reference's ``output[0]``).
    """

    EOS_ID = 49407  # hard-coded at text_encoder.py:71

    def __init__(self, cfg: ClipTextConfig):
        super().__init__()
        self.config = cfg
        self.text_model = ClipTextTransformer(cfg)
        self.register_buffer("null_embedding",
                             torch.zeros(cfg.max_position_embeddings, cfg.hidden_size))
        self._use_fixed_special_embedding = False

    # text_encoder.py:28-32
    def set_null_embedding(self, null_embedding: torch.Tensor):
        self.null_embedding = null_embedding
        self._use_fixed_special_embedding = True

    def get_input_embeddings(self):
        return self.text_model.embeddings.token_embedding

    def resize_token_embeddings(self, n: int):
        old = self.text_model.embeddings.token_embedding
        if n == old.num_embeddings:
            return old
        new = nn.Embedding(n, old.embedding_dim).to(old.weight)
        with torch.no_grad():
            k = min(n, old.num_embeddings)
            new.weight[:k] = old.weight[:k]
        self.text_model.embeddings.token_embedding = new
        return new

    # train_textboost.py:700-710 (peft freezes every non-LoRA parameter inside add_adapter)
    def add_adapter(self, r: int = 4, lora_alpha: int | None = None,
                    target_modules: Sequence[str] = ("q_proj", "k_proj", "v_proj")):
        lora_alpha = r if lora_alpha is None else lora_alpha
        for p in self.parameters():
            p.requires_grad_(False)
        for layer in self.text_model.encoder.layers:
            for name in target_modules:
                holder = layer.self_attn if hasattr(layer.self_attn, name) else layer.mlp
                setattr(holder, name, LoraLinear(getattr(holder, name), r, lora_alpha))

    def forward(self, input_ids):
        out = self.text_model(input_ids)
        null_pos = input_ids[:, 1] == self.EOS_ID  # text_encoder.py:71
        if null_pos.any():
            out = out.clone()
            out[null_pos] = self.null_embedding.to(out.dtype).unsqueeze(0).repeat(
                int(null_pos.sum()), 1, 1)
        if self._use_fixed_special_embedding:  # text_encoder.py:81-86
            out = out.clone()
            out[:, 0] = self.null_embedding[0].to(out.dtype).unsqueeze(0).repeat(out.shape[0], 1)
        return out


def init_clip_(model: TextBoostModelRef, seed: int = 0, std: float = 0.02):
    """Deterministic random init (no checkpoints offline): N(0, std) weights as in CLIP's
    initializer_range, LayerNorm weight ~ 1 + N(0, .1) so affine terms are exercised."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if "layer_norm" in name and name.endswith("weight"):
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            elif name.endswith("bias"):
                p.copy_(0.02 * torch.randn(p.shape, generator=g))
            elif "fc2" in name or "out_proj" in name:
                p.copy_(torch.randn(p.shape, generator=g) * std / math.sqrt(2.0))
            else:
                p.copy_(torch.randn(p.shape, generator=g) * std)
    return model
