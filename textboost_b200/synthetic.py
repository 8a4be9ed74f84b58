"""Synthetic workload of BASELINE.json: random-init weights at the exact SD-1.5 / SD-2.1 shapes,
synthetic latents and literal prompt ids.  No checkpoints, datasets or tokenizer files exist offline
(SURVEY.md §8d), so this is what bench.py, smoke() and the GPU tests run on.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch

from .precision import POLICY
from .clip import ClipConfig, ClipEngine
from .trainer import TextBoostTrainer
from .unet import UNetConfig, UNetEngine

BOS, EOS = 49406, 49407


# ---------------------------------------------------------------------------------- shapes
def unet_shapes(cfg: UNetConfig) -> Dict[str, Tuple[int, ...]]:
    """diffusers UNet2DConditionModel state-dict keys -> shapes."""
    ch = cfg.block_out_channels
    T = ch[0] * 4
    ctx = cfg.cross_attention_dim
    s: Dict[str, Tuple[int, ...]] = {}

    def conv(p, co, ci, k):
        s[p + ".weight"] = (co, ci, k, k)
        s[p + ".bias"] = (co,)

    def lin(p, co, ci, bias=True):
        s[p + ".weight"] = (co, ci)
        if bias:
            s[p + ".bias"] = (co,)

    def norm(p, c):
        s[p + ".weight"] = (c,)
        s[p + ".bias"] = (c,)

    def resnet(p, ci, co):
        norm(p + "norm1", ci)
        conv(p + "conv1", co, ci, 3)
        lin(p + "time_emb_proj", co, T)
        norm(p + "norm2", co)
        conv(p + "conv2", co, co, 3)
        if ci != co:
            conv(p + "conv_shortcut", co, ci, 1)

    def transformer(p, c):
        norm(p + "norm", c)
        if cfg.use_linear_projection:
            lin(p + "proj_in", c, c)
            lin(p + "proj_out", c, c)
        else:
            conv(p + "proj_in", c, c, 1)
            conv(p + "proj_out", c, c, 1)
        t = p + "transformer_blocks.0."
        for n in ("norm1", "norm2", "norm3"):
            norm(t + n, c)
        for a, kd in (("attn1", c), ("attn2", ctx)):
            lin(t + a + ".to_q", c, c, False)
            lin(t + a + ".to_k", c, kd, False)
            lin(t + a + ".to_v", c, kd, False)
            lin(t + a + ".to_out.0", c, c)
        lin(t + "ff.net.0.proj", 8 * c, c)
        lin(t + "ff.net.2", c, 4 * c)

    conv("conv_in", ch[0], cfg.in_channels, 3)
    lin("time_embedding.linear_1", T, ch[0])
    lin("time_embedding.linear_2", T, T)
    out = ch[0]
    for i, c in enumerate(ch):
        cin, out = out, c
        for j in range(cfg.layers_per_block):
            resnet(f"down_blocks.{i}.resnets.{j}.", cin if j == 0 else out, out)
            if cfg.down_has_attn[i]:
                transformer(f"down_blocks.{i}.attentions.{j}.", out)
        if i != len(ch) - 1:
            conv(f"down_blocks.{i}.downsamplers.0.conv", out, out, 3)
    resnet("mid_block.resnets.0.", ch[-1], ch[-1])
    transformer("mid_block.attentions.0.", ch[-1])
    resnet("mid_block.resnets.1.", ch[-1], ch[-1])
    rev = list(reversed(ch))
    rev_attn = list(reversed(cfg.down_has_attn))
    out = rev[0]
    n = cfg.layers_per_block + 1
    for i, c in enumerate(rev):
        prev, out = out, c
        cin = rev[min(i + 1, len(ch) - 1)]
        for j in range(n):
            skip = cin if j == n - 1 else out
            rin = prev if j == 0 else out
            resnet(f"up_blocks.{i}.resnets.{j}.", rin + skip, out)
            if rev_attn[i]:
                transformer(f"up_blocks.{i}.attentions.{j}.", out)
        if i != len(ch) - 1:
            conv(f"up_blocks.{i}.upsamplers.0.conv", out, out, 3)
    norm("conv_norm_out", ch[0])
    conv("conv_out", cfg.out_channels, ch[0], 3)
    return s


def clip_shapes(cfg: ClipConfig, vocab: int) -> Dict[str, Tuple[int, ...]]:
    D, F = cfg.hidden_size, cfg.intermediate_size
    s = {"text_model.embeddings.token_embedding.weight": (vocab, D),
         "text_model.embeddings.position_embedding.weight": (cfg.max_position_embeddings, D)}
    for l in range(cfg.num_hidden_layers):
        p = f"text_model.encoder.layers.{l}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            s[p + f"self_attn.{n}.weight"] = (D, D)
            s[p + f"self_attn.{n}.bias"] = (D,)
        for n in ("layer_norm1", "layer_norm2"):
            s[p + n + ".weight"] = (D,)
            s[p + n + ".bias"] = (D,)
        s[p + "mlp.fc1.weight"] = (F, D)
        s[p + "mlp.fc1.bias"] = (F,)
        s[p + "mlp.fc2.weight"] = (D, F)
        s[p + "mlp.fc2.bias"] = (D,)
    s["text_model.final_layer_norm.weight"] = (D,)
    s["text_model.final_layer_norm.bias"] = (D,)
    return s


# ---------------------------------------------------------------------------------- random init
def random_unet_sd(cfg: UNetConfig, device, seed=0, dtype=None):
    """N(0, 1/fan_in) conv/linear weights (residual-branch outputs damped), norm affine 1 + N(0,.1):
    activations stay O(1) through the ~60 layers (a plain N(0, .02) init collapses the GroupNorm inputs)."""
    g = torch.Generator(device=device).manual_seed(seed)
    sd = {}
    for k, shp in unet_shapes(cfg).items():
        if len(shp) == 1:
            if "norm" in k and k.endswith("weight"):
                t = 1.0 + 0.1 * torch.randn(shp, generator=g, device=device)
            else:
                t = 0.05 * torch.randn(shp, generator=g, device=device)
        else:
            fan_in = math.prod(shp[1:])
            gain = 0.5 if any(x in k for x in ("conv2.", "to_out.0", "ff.net.2", "proj_out")) else 1.0
            t = torch.randn(shp, generator=g, device=device) * (gain / math.sqrt(fan_in))
        sd[k] = t.to(dtype or POLICY.act)
    return sd


def random_vae_sd(cfg=None, device="cpu", seed=0, dtype=torch.float32, with_decoder=False):
    """Random AutoencoderKL encoder weights (fp32 on disk, as SD checkpoints ship them): N(0, 1/fan_in) convs /
    linears, norm affine 1 + N(0,.1), small biases."""
    from . import vae
    cfg = cfg or vae.VAEConfig()
    g = torch.Generator(device=device).manual_seed(seed)
    sd = {}
    shapes = dict(vae.vae_encoder_shapes(cfg))
    if with_decoder:
        shapes.update(vae.vae_decoder_shapes(cfg))
    for k, shp in shapes.items():
        if len(shp) == 1:
            base = 1.0 if ("norm" in k and k.endswith("weight")) else 0.0
            t = base + (0.1 if base else 0.05) * torch.randn(shp, generator=g, device=device)
        else:
            t = torch.randn(shp, generator=g, device=device) / math.sqrt(math.prod(shp[1:]))
        sd[k] = t.to(dtype)
    return sd


def random_clip_sd(cfg: ClipConfig, vocab: int, device, seed=0):
    g = torch.Generator(device=device).manual_seed(seed)
    sd = {}
    for k, shp in clip_shapes(cfg, vocab).items():
        if "layer_norm" in k and k.endswith("weight"):
            t = 1.0 + 0.1 * torch.randn(shp, generator=g, device=device)
        elif k.endswith("bias"):
            t = 0.02 * torch.randn(shp, generator=g, device=device)
        else:
            std = 0.02 / math.sqrt(2.0) if ("fc2" in k or "out_proj" in k) else 0.02
            t = torch.randn(shp, generator=g, device=device) * std
        sd[k] = t
    return sd


# ---------------------------------------------------------------------------------- inputs
def instance_ids(B: int, placeholder_id: int, L: int = 77) -> torch.Tensor:
    """'a <dog> dog' -> [BOS, 320, <placeholder>, 1929, EOS x 73] (CLIP BPE ids; structure is what matters)."""
    row = torch.full((L,), EOS, dtype=torch.int64)
    row[0], row[1], row[2], row[3] = BOS, 320, placeholder_id, 1929
    return row.unsqueeze(0).repeat(B, 1)


def prior_ids(B: int, seed: int, L: int = 77, null_prob: float = 0.1) -> torch.Tensor:
    """[BOS, r_1..r_k, EOS...], r ~ U{1000..40000}, k ~ U{3..20}; ~10 % empty prompts (text_encoder.py:71)."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.full((B, L), EOS, dtype=torch.int64)
    ids[:, 0] = BOS
    for b in range(B):
        if torch.rand((), generator=g).item() < null_prob:
            continue
        k = int(torch.randint(3, 21, (), generator=g))
        ids[b, 1:1 + k] = torch.randint(1000, 40001, (k,), generator=g)
    return ids


def batch(B: int, H: int, seed: int, placeholder_id: int, device, rank: int = 0):
    g = torch.Generator().manual_seed(seed + rank)
    lat = torch.randn(B, 4, H, H, generator=g)
    noise = torch.randn(B, 4, H, H, generator=g)
    t = torch.randint(0, 1000, (B,), generator=g)
    return {"latents": lat.to(device), "noise": noise.to(device), "timesteps": t.to(device),
            "input_ids": instance_ids(B, placeholder_id).to(device),
            "prior_ids": prior_ids(B, seed + 1000 + rank).to(device)}


# ---------------------------------------------------------------------------------- whole trainer
def model_configs(model: str):
    """'sd15' (config 2/3/4), 'sd21' (config 5: SD-2.x widths + OpenCLIP-H) or 'tiny' (same topology,
    small widths: smoke() and the oracle-speed parity tests)."""
    if model == "sd15":
        return UNetConfig.sd15(), ClipConfig.clip_l()
    if model == "sd21":
        return UNetConfig.sd21(), ClipConfig.openclip_h()
    if model == "tiny":
        return (UNetConfig(block_out_channels=(64, 128, 128, 128), attention_head_dim=(2, 2, 2, 2),
                           cross_attention_dim=128, sample_size=16),
                ClipConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2))
    raise ValueError(model)


def build_trainer(model: str = "sd15", device="cuda", seed: int = 42, n_added: int = 1, lora_r: int = 4,
                  kpl_weight: float = 0.1, lora_b_std: float = 0.0, keep_sd: bool = False,
                  lora_targets=("q_proj", "k_proj", "v_proj"), lora_alpha=None, unet_lora_r: int = 0,
                  **trainer_kw) -> TextBoostTrainer:
    """Random-init TextBoost trainer.  n_added rows are appended to the vocabulary and initialised from an
    existing row (utils.add_token, textboost/utils.py:117-166).  keep_sd=True stashes the generated
    state dicts on ``trainer.synthetic`` so a checker can rebuild the same model elsewhere."""
    ucfg, ccfg = model_configs(model)
    usd = random_unet_sd(ucfg, device, seed)
    unet = UNetEngine(ucfg, usd)
    if unet_lora_r:  # --unet_params_to_train crossattn_kv (bf16 policy only, see TextBoostTrainer)
        kvl = unet.add_cross_kv_lora(unet_lora_r, seed=seed + 5)
        if lora_b_std > 0:
            g = torch.Generator(device=device).manual_seed(seed + 6)
            kvl.B().copy_(lora_b_std * torch.randn(kvl.B().shape, generator=g, device=device))
    csd = random_clip_sd(ccfg, ccfg.vocab_size, device, seed + 1)
    null = torch.randn((ccfg.max_position_embeddings, ccfg.hidden_size),
                       generator=torch.Generator().manual_seed(seed + 2))
    te0 = ClipEngine(ccfg, csd, device, lora_r=0) if kpl_weight > 0 else None
    if te0 is not None:
        te0.set_null_embedding(null)
    emb = csd["text_model.embeddings.token_embedding.weight"]
    csd_t = dict(csd)
    csd_t["text_model.embeddings.token_embedding.weight"] = torch.cat([emb, emb[1929:1929 + n_added]], 0)
    te = ClipEngine(ccfg, csd_t, device, lora_r=lora_r, lora_alpha=lora_alpha, n_base=ccfg.vocab_size, seed=seed + 3,
                    lora_targets=lora_targets)
    te.set_null_embedding(null)
    if lora_b_std > 0:  # step-0 dL/dA is exactly 0 with B = 0 (SURVEY.md trap 13): tests use B != 0
        g = torch.Generator(device=device).manual_seed(seed + 4)
        te.state.b_segment().copy_(lora_b_std * torch.randn(te.state.n_b, generator=g, device=device))
    tr = TextBoostTrainer(unet, te, te0, kpl_weight=kpl_weight, **trainer_kw)
    tr.synthetic = {"model": model, "unet_cfg": ucfg, "clip_cfg": ccfg, "n_added": n_added, "lora_r": lora_r,
                    "null": null}
    if keep_sd:
        tr.synthetic.update(unet_sd=usd, clip_sd=csd)
    return tr


# ---------------------------------------------------------------------------------- offline stand-ins
class LiteralTokenizer:
    """Duck-typed tokenizer for a box with no vocab.json / merges.txt: whitespace words map to stable ids in
    [1000, 40000]; BOS/EOS are CLIP's.  Implements exactly what textboost.utils.add_token uses
    (encode / add_tokens / convert_tokens_to_ids / __len__) plus __call__ -> padded [1, L] ids."""

    def __init__(self, vocab_size: int = 49408, model_max_length: int = 77):
        self.base, self.model_max_length = vocab_size, model_max_length
        self.added: Dict[str, int] = {}

    def __len__(self):
        return self.base + len(self.added)

    def _word(self, w: str) -> int:
        if w in self.added:
            return self.added[w]
        h = 0
        for c in w.encode():
            h = (h * 131 + c) % 39001
        return 1000 + h

    def encode(self, text: str, add_special_tokens: bool = True):
        ids = [self._word(w) for w in text.split()]
        return [BOS] + ids + [EOS] if add_special_tokens else ids

    def add_tokens(self, tokens) -> int:
        n = 0
        for t in ([tokens] if isinstance(tokens, str) else tokens):
            if t not in self.added:
                self.added[t] = len(self)
                n += 1
        return n

    def convert_tokens_to_ids(self, tokens):
        if isinstance(tokens, str):
            return self._word(tokens)
        return [self._word(t) for t in tokens]

    def __call__(self, text: str, truncation=True, padding="max_length", max_length=None, return_tensors="pt"):
        """The call textboost.dataset.tokenize_prompt makes: padded [1, L] ids (CLIP pads with EOS) + mask."""
        from types import SimpleNamespace
        L = max_length or self.model_max_length
        ids = self.encode(text)
        if len(ids) > L:
            ids = ids[:L - 1] + [EOS]
        row = torch.full((1, L), EOS, dtype=torch.int64)
        row[0, :len(ids)] = torch.tensor(ids)
        mask = torch.zeros((1, L), dtype=torch.int64)
        mask[0, :len(ids)] = 1
        return SimpleNamespace(input_ids=row, attention_mask=mask)


LITERAL_MARKER = "literal_tokenizer.json"


def write_literal_tokenizer_marker(directory: str, vocab_size: int = 49408):
    """tokenizer/literal_tokenizer.json: the explicit opt-in that lets load_tokenizer hand out the LiteralTokenizer
    for a synthetic checkpoint.  A real checkpoint without vocab.json / merges.txt raises instead of silently
    training on hashed ids."""
    import json
    import os
    os.makedirs(os.path.join(directory, "tokenizer"), exist_ok=True)
    with open(os.path.join(directory, "tokenizer", LITERAL_MARKER), "w") as f:
        json.dump({"tokenizer_class": "LiteralTokenizer", "vocab_size": vocab_size, "model_max_length": 77}, f)


def load_tokenizer(name_or_dir: str, allow_literal: bool = False):
    """The reference's ``AutoTokenizer.from_pretrained(..., subfolder="tokenizer", use_fast=False)``
    (/root/reference/train_textboost.py:630-638) for a tokenizer directory or hub id.

    * directory with vocab.json + merges.txt -> transformers' CLIP tokenizer;
    * directory holding the literal marker written by write_pretrained (a synthetic checkpoint), or
      allow_literal=True (--synthetic_data) -> LiteralTokenizer;
    * a name that is not a local directory -> AutoTokenizer.from_pretrained(name) (needs the hub / a cache);
    * anything else -> OSError.  Never a silent fallback: hashed ids on a real checkpoint would train garbage."""
    import json
    import os
    if os.path.isdir(name_or_dir):
        if os.path.exists(os.path.join(name_or_dir, "vocab.json")):
            if not os.path.exists(os.path.join(name_or_dir, "merges.txt")):
                raise OSError(f"{name_or_dir}: vocab.json without merges.txt")
            from transformers import AutoTokenizer
            return AutoTokenizer.from_pretrained(name_or_dir, use_fast=False)
        marker = os.path.join(name_or_dir, LITERAL_MARKER)
        if os.path.exists(marker):
            with open(marker) as f:
                m = json.load(f)
            return LiteralTokenizer(m.get("vocab_size", 49408), m.get("model_max_length", 77))
        if allow_literal:
            return LiteralTokenizer()
        raise OSError(f"{name_or_dir}: no vocab.json / merges.txt (and no {LITERAL_MARKER} marker of a synthetic "
                      "checkpoint); refusing to fall back to the literal stand-in tokenizer")
    if allow_literal:  # --synthetic_data: the explicit opt-in
        return LiteralTokenizer()
    if os.path.isdir(os.path.dirname(os.path.abspath(name_or_dir))) and os.sep in name_or_dir:
        raise OSError(f"{name_or_dir}: tokenizer directory not found")  # <checkpoint>/tokenizer is missing
    from transformers import AutoTokenizer
    return AutoTokenizer.from_pretrained(name_or_dir, use_fast=False)


def write_pretrained(directory: str, model: str = "tiny", seed: int = 0, prediction_type: str = "epsilon",
                     vae_channels=None):
    """Write a random-init checkpoint in the diffusers directory layout (unet/, text_encoder/, scheduler/, and
    vae/ when `vae_channels` = its block_out_channels is given) that train_textboost.py's from_pretrained calls
    read (train_textboost.py:633-656)."""
    import json
    import os
    from safetensors.torch import save_file
    from . import text_encoder as te_mod
    from . import unet_model
    ucfg, ccfg = model_configs(model)
    usd = {k: v.contiguous() for k, v in random_unet_sd(ucfg, "cpu", seed).items()}
    csd = {k: v.contiguous() for k, v in random_clip_sd(ccfg, ccfg.vocab_size, "cpu", seed + 1).items()}
    os.makedirs(os.path.join(directory, "unet"), exist_ok=True)
    os.makedirs(os.path.join(directory, "text_encoder"), exist_ok=True)
    os.makedirs(os.path.join(directory, "scheduler"), exist_ok=True)
    save_file(usd, os.path.join(directory, "unet", "diffusion_pytorch_model.safetensors"), metadata={"format": "pt"})
    with open(os.path.join(directory, "unet", "config.json"), "w") as f:
        json.dump(unet_model.config_to_dict(ucfg), f, indent=2)
    save_file(csd, os.path.join(directory, "text_encoder", "model.safetensors"), metadata={"format": "pt"})
    with open(os.path.join(directory, "text_encoder", "config.json"), "w") as f:
        json.dump({**te_mod.config_to_dict(ccfg), "architectures": ["CLIPTextModel"],
                   "model_type": "clip_text_model"}, f, indent=2)
    if vae_channels:  # vae/ in the diffusers layout (fp32 weights, as SD checkpoints ship them)
        from . import vae
        vcfg = vae.VAEConfig(block_out_channels=tuple(vae_channels))
        os.makedirs(os.path.join(directory, "vae"), exist_ok=True)
        vsd = {k: v.contiguous() for k, v in random_vae_sd(vcfg, "cpu", seed + 2, with_decoder=True).items()}
        save_file(vsd, os.path.join(directory, "vae", "diffusion_pytorch_model.safetensors"),
                  metadata={"format": "pt"})
        with open(os.path.join(directory, "vae", "config.json"), "w") as f:
            json.dump(vae.config_to_dict(vcfg), f, indent=2)
    write_literal_tokenizer_marker(directory)
    with open(os.path.join(directory, "scheduler", "scheduler_config.json"), "w") as f:
        json.dump({"_class_name": "DDPMScheduler", "num_train_timesteps": 1000, "beta_start": 0.00085,
                   "beta_end": 0.012, "beta_schedule": "scaled_linear", "prediction_type": prediction_type}, f)
    return usd, csd
