"""``TextBoostModel`` — host-side mirror of /root/reference/textboost/text_encoder.py:17-87 over the B200
CLIP engine (textboost_b200.clip.ClipEngine -> libtextboost_b200.so).

Same surface as the reference class on the TextBoost path:

  TextBoostModel.from_pretrained(path, subfolder="text_encoder", revision, variant)   train_textboost.py:646
  .set_null_embedding(path | tensor)                                                  text_encoder.py:28-32
  copy.deepcopy(model).eval().requires_grad_(False)                                   train_textboost.py:650
  .get_input_embeddings().weight / .resize_token_embeddings(n)                        utils.py:158-165
  .text_model.encoder.requires_grad_(True) / .parameters()                            train_textboost.py:701, 835
  .add_adapter(LoraConfig) / .load_adapter(dir, name) / .set_adapter(name)            :709, inference.py:56-58
  .named_parameters() with "lora" / "lora_B" / "token_embedding" in the names          :727-733, :1120-1126
  .to(device[, dtype]) / .device / .dtype / .config                                   :810, :919-939
  model(input_ids, attention_mask=None, ..., return_dict=False)[0]                    utils.py:18-24
  .save_pretrained(dir)  (adapter-only once an adapter is attached)                   :1178-1182, :1241-1243

The weights live on the host in the HF key layout until the model is moved to a CUDA device; ``.to("cuda")``
builds the engine (fp16 GEMM operands, fp32 master / residual stream, LoRA fused into the QKV GEMM).  After
that the trainable tensors the reference touches — every ``lora_A`` / ``lora_B`` and the added token rows —
are VIEWS of the engine's flat parameter / gradient buffers, so the all-reduce and the fused AdamW see
exactly what ``named_parameters()`` reports.  ``forward`` is a ``torch.autograd.Function``: ``loss.backward()``
runs the hand-derived CLIP backward in the library and deposits into those gradient views.

There is no CPU execution path: calling the model before it is on an sm_100 device raises.
"""
from __future__ import annotations

import copy
import json
import os
from types import SimpleNamespace
from typing import Dict, Iterator, Optional, Tuple

import torch

from .precision import POLICY
from .clip import EOS_ID, LORA_TARGETS, ClipConfig, ClipEngine, canonical_targets
from .lora import LoraConfig

F32 = torch.float32

_CFG_KEYS = ("vocab_size", "hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads",
             "max_position_embeddings", "hidden_act", "layer_norm_eps")


class ModelOutput(tuple):
    """(last_hidden_state, pooler_output) that also answers to the attribute names transformers uses."""

    def __new__(cls, last_hidden_state, pooler_output):
        return super().__new__(cls, (last_hidden_state, pooler_output))

    last_hidden_state = property(lambda self: self[0])
    pooler_output = property(lambda self: self[1])


class _ClipFunction(torch.autograd.Function):
    """Engine forward / backward behind autograd.  `anchor` is a 1-element tensor that requires grad, so the
    graph reaches this node; the real gradients are accumulated by the library into engine.state.grads."""

    @staticmethod
    def forward(ctx, anchor, engine, input_ids):
        ctx.engine = engine
        out = engine.forward(input_ids, save_for_backward=True)
        ctx.saved = engine.pop_ctx()
        return out

    @staticmethod
    def backward(ctx, d_out):
        ctx.engine.backward(d_out.to(F32).contiguous().clone(), ctx=ctx.saved)
        ctx.saved = None
        return torch.zeros(1, device=d_out.device, dtype=F32), None, None


class _TokenEmbedding:
    """What ``get_input_embeddings()`` returns: ``.weight`` [V, D] fp32 and ``requires_grad_``."""

    def __init__(self, owner: "TextBoostModel"):
        self._o = owner

    @property
    def weight(self) -> torch.Tensor:
        o = self._o
        if o._engine is None:
            return o._sd["text_model.embeddings.token_embedding.weight"]
        e = o._engine  # dense view of the lazily decayed matrix: base*c ‖ added rows
        return torch.cat([e.tok_base * e.decay, e.state.rows()], 0)

    @property
    def num_embeddings(self) -> int:
        return self._o.vocab_size

    def requires_grad_(self, flag: bool = True):
        self._o._train_embedding = bool(flag)
        return self

    def parameters(self):
        yield self.weight


class _Encoder:
    """``text_encoder.text_model.encoder``: the handle train_textboost.py:701/835/1131 use."""

    def __init__(self, owner: "TextBoostModel"):
        self._o = owner

    def requires_grad_(self, flag: bool = True):
        self._o._encoder_requires_grad = bool(flag)
        return self

    def named_parameters(self):
        for k, v in self._o.named_parameters():
            if k.startswith("text_model.encoder."):
                yield k[len("text_model.encoder."):], v

    def parameters(self):
        """The reference builds its second optimiser group from the requires_grad members of this
        (train_textboost.py:835): after peft's injection those are exactly the LoRA tensors."""
        for _, v in self.named_parameters():
            yield v


class TextBoostModel:
    config_class = ClipConfig
    _null_override = True

    def __init__(self, config: ClipConfig, state_dict: Optional[Dict[str, torch.Tensor]] = None):
        self.config = config if isinstance(config, ClipConfig) else ClipConfig(**config)
        self.config.use_return_dict = True
        D = self.config.hidden_size
        self._sd: Dict[str, torch.Tensor] = {}
        if state_dict is not None:
            for k, v in state_dict.items():
                if k == "null_embedding":
                    continue
                self._sd[k] = v.detach().to("cpu", F32).clone()
        # registered buffer in the reference (text_encoder.py:19-25): zeros until set_null_embedding
        self.null_embedding = torch.zeros(self.config.max_position_embeddings, D, dtype=F32)
        if state_dict is not None and "null_embedding" in state_dict:
            self.null_embedding = state_dict["null_embedding"].detach().to("cpu", F32).clone()
        self._use_fixed_special_embedding = False
        self._n_base = self.config.vocab_size  # ids >= this were appended by add_token (utils.py:117-166)
        self._lora: Optional[LoraConfig] = None
        self._engine: Optional[ClipEngine] = None
        self._device = torch.device("cpu")
        self._dtype = F32
        self._train_embedding = False
        self._encoder_requires_grad = False
        self._requires_grad = True
        self.training = True
        self.text_model = SimpleNamespace(encoder=_Encoder(self), embeddings=SimpleNamespace(
            token_embedding=_TokenEmbedding(self)))
        self._anchor = None
        self._seed = 0

    # ------------------------------------------------------------------ load / save
    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder: Optional[str] = None, revision=None,
                        variant: Optional[str] = None, **_unused):
        """Reads ``config.json`` + ``model[.variant].safetensors`` (or ``pytorch_model.bin``) in the
        transformers layout.  A missing ``null_embedding`` key is tolerated: stock checkpoints lack it."""
        d = os.path.join(pretrained_model_name_or_path, subfolder) if subfolder else pretrained_model_name_or_path
        with open(os.path.join(d, "config.json")) as f:
            raw = json.load(f)
        cfg = ClipConfig(**{k: raw[k] for k in _CFG_KEYS if k in raw})
        stem = "model" + (f".{variant}" if variant else "")
        st_path = os.path.join(d, stem + ".safetensors")
        if os.path.exists(st_path):
            from safetensors.torch import load_file
            sd = load_file(st_path)
        else:
            bin_path = os.path.join(d, "pytorch_model" + (f".{variant}" if variant else "") + ".bin")
            if not os.path.exists(bin_path):
                raise OSError(f"no {stem}.safetensors or pytorch_model.bin under {d}")
            sd = torch.load(bin_path, map_location="cpu", weights_only=True)
        sd.pop("text_model.embeddings.position_ids", None)
        m = cls(cfg, sd)
        m._raw_config = raw
        return m

    def save_pretrained(self, save_directory: str, safe_serialization: bool = True, **_unused):
        """With an adapter attached this writes the adapter only — ``adapter_config.json`` +
        ``adapter_model.safetensors`` with ``base_model.model.`` prefixed, ``.default``-stripped keys — which
        is what transformers' PeftAdapterMixin does for the reference (train_textboost.py:1181, 1243) and
        what ``inference.py:56-58`` loads.  Without an adapter: the full model in the transformers layout."""
        os.makedirs(save_directory, exist_ok=True)
        if self._lora is not None:
            self._lora.save(save_directory)
            out = {}
            for k, v in self.named_parameters():
                if "lora_" in k:
                    out["base_model.model." + k.replace(".default", "")] = v.detach().cpu().contiguous().clone()
            _save_tensors(out, save_directory, "adapter_model", safe_serialization)
            return
        sd = {k: v.detach().cpu().contiguous() for k, v in self.state_dict().items()}
        _save_tensors(sd, save_directory, "model" if safe_serialization else "pytorch_model", safe_serialization)
        raw = dict(getattr(self, "_raw_config", {}))
        raw.update({k: getattr(self.config, k) for k in _CFG_KEYS})
        raw["vocab_size"] = self.vocab_size
        raw.setdefault("architectures", ["CLIPTextModel"])
        raw.setdefault("model_type", "clip_text_model")
        with open(os.path.join(save_directory, "config.json"), "w") as f:
            json.dump(raw, f, indent=2, sort_keys=True)

    # ------------------------------------------------------------------ reference API
    def set_null_embedding(self, null_embedding):
        """text_encoder.py:28-32 (a str is a ``torch.save``d tensor).  The SD-2.1 asset is [77,1024]; a width
        mismatch — the reference's crash on SD-1.x, SURVEY.md D6 — is reported here instead of at forward."""
        if isinstance(null_embedding, str):
            null_embedding = torch.load(null_embedding, map_location="cpu", weights_only=True)
        want = (self.config.max_position_embeddings, self.config.hidden_size)
        if tuple(null_embedding.shape) != want:
            raise RuntimeError(f"null_embedding has shape {tuple(null_embedding.shape)}, this text encoder "
                               f"needs {want}")
        self.null_embedding = null_embedding.detach().to("cpu", F32).clone()
        self._use_fixed_special_embedding = True
        if self._engine is not None:
            self._engine.set_null_embedding(self.null_embedding)

    @property
    def vocab_size(self) -> int:
        if self._engine is not None:
            return self._engine.n_base + self._engine.state.n_rows
        return self._sd["text_model.embeddings.token_embedding.weight"].shape[0]

    def get_input_embeddings(self) -> _TokenEmbedding:
        return self.text_model.embeddings.token_embedding

    def resize_token_embeddings(self, new_num_tokens: int):
        """Grow (never shrink) the token matrix; new rows start as N(0, 0.02) like transformers' _init_weights
        and are overwritten by add_token (utils.py:161-165)."""
        if self._engine is not None:
            raise RuntimeError("resize_token_embeddings must precede .to(cuda): added rows are laid out in the "
                               "engine's flat trainable buffer when it is built")
        k = "text_model.embeddings.token_embedding.weight"
        w = self._sd[k]
        if new_num_tokens < w.shape[0]:
            raise ValueError("shrinking the vocabulary is not part of the TextBoost path")
        if new_num_tokens > w.shape[0]:
            g = torch.Generator().manual_seed(self._seed + w.shape[0])
            extra = torch.randn((new_num_tokens - w.shape[0], w.shape[1]), generator=g) * 0.02
            self._sd[k] = torch.cat([w, extra], 0)
        return self.get_input_embeddings()

    def add_adapter(self, adapter_config: LoraConfig, adapter_name: str = "default"):
        """peft injection (train_textboost.py:702-709): A ~ N(0, (1/r)^2), B = 0, everything that is not a
        LoRA tensor stops requiring grad (peft's mark_only_adapters_as_trainable)."""
        if self._engine is not None:
            raise RuntimeError("add_adapter must precede .to(cuda)")
        if self._lora is not None:
            raise ValueError(f"Adapter with name {adapter_name} already exists. Please use a different name.")
        adapter_config.validate()
        self._lora = adapter_config
        r, D = adapter_config.r, self.config.hidden_size
        g = torch.Generator().manual_seed(self._seed + 7)
        for l in range(self.config.num_hidden_layers):
            for t in canonical_targets(adapter_config.target_modules):
                p = f"text_model.encoder.layers.{l}.self_attn.{t}."
                for kind in ("weight", "bias"):
                    self._sd[p + "base_layer." + kind] = self._sd.pop(p + kind)
                if adapter_config.init_lora_weights == "gaussian":
                    a = torch.randn((r, D), generator=g) / r
                else:  # peft default: kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(D), 1/sqrt(D))
                    a = (torch.rand((r, D), generator=g) * 2 - 1) / D ** 0.5
                self._sd[p + "lora_A.default.weight"] = a
                self._sd[p + "lora_B.default.weight"] = torch.zeros((D, r))
        self._encoder_requires_grad = False
        self._train_embedding = False

    def load_adapter(self, peft_model_id: str, adapter_name: str = "default", **_unused):
        """inference.py:55-58: attach the adapter saved by ``save_pretrained`` above."""
        cfg = LoraConfig.load(peft_model_id)
        st = os.path.join(peft_model_id, "adapter_model.safetensors")
        if os.path.exists(st):
            from safetensors.torch import load_file
            sd = load_file(st)
        else:
            sd = torch.load(os.path.join(peft_model_id, "adapter_model.bin"), map_location="cpu", weights_only=True)
        if self._lora is None:
            self.add_adapter(cfg, adapter_name)
        for k, v in sd.items():
            k = k[len("base_model.model."):] if k.startswith("base_model.model.") else k
            k = k.replace("lora_A.weight", "lora_A.default.weight").replace("lora_B.weight", "lora_B.default.weight")
            if k not in self._sd:
                raise KeyError(f"unexpected adapter key {k}")
            self._sd[k] = v.detach().to("cpu", F32).clone()

    def set_adapter(self, adapter_name):  # one adapter only
        return None

    # ------------------------------------------------------------------ nn.Module-like plumbing
    def eval(self):
        self.training = False
        return self

    def train(self, mode: bool = True):
        self.training = mode
        return self

    def requires_grad_(self, flag: bool = True):
        self._requires_grad = self._train_embedding = self._encoder_requires_grad = bool(flag)
        return self

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
            self._dtype = dtype  # recorded only: GEMM operands are fp16, master / residual stream fp32
        if device is not None:
            device = torch.device(device)
            if device.type == "cuda" and self._engine is None:
                self._materialize(device)
            elif device.type != "cuda" and self._engine is not None:
                raise RuntimeError("the B200 text encoder cannot be moved off the GPU (no CPU path)")
            self._device = device
        return self

    def cuda(self, device=None):
        return self.to(torch.device("cuda", torch.cuda.current_device() if device is None else device))

    def __deepcopy__(self, memo):
        new = TextBoostModel.__new__(TextBoostModel)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k == "text_model":
                continue
            setattr(new, k, copy.deepcopy(v, memo))
        new.text_model = SimpleNamespace(encoder=_Encoder(new), embeddings=SimpleNamespace(
            token_embedding=_TokenEmbedding(new)))
        return new

    def _materialize(self, device):
        r = self._lora.r if self._lora is not None else 0
        alpha = self._lora.lora_alpha if self._lora is not None else None
        targets = self._lora.target_modules if self._lora is not None else LORA_TARGETS
        self._engine = ClipEngine(self.config, self._sd, device, lora_r=r, lora_alpha=alpha,
                                  n_base=self._n_base, seed=self._seed, lora_targets=targets)
        self._engine.null_embedding = self.null_embedding.to(device)
        self._engine.use_fixed_special = self._use_fixed_special_embedding
        self._engine.null_override = self._null_override
        self._anchor = torch.zeros(1, device=device, dtype=F32, requires_grad=True)
        # host copies of what now lives in the engine's flat buffer would go stale: drop them
        for k in [k for k in self._sd if "lora_" in k]:
            del self._sd[k]
        k = "text_model.embeddings.token_embedding.weight"
        self._sd[k] = self._sd[k][:0]

    @property
    def engine(self) -> ClipEngine:
        if self._engine is None:
            raise RuntimeError("TextBoostModel has no CPU execution path: move it to an sm_100 device first "
                               "(.to('cuda')); the CUDA extension is the product")
        return self._engine

    # ------------------------------------------------------------------ parameters
    def _trainable_views(self) -> Iterator[Tuple[str, torch.Tensor, torch.Tensor]]:
        e = self._engine
        st = e.state
        r = st.r
        for l in range(e.nl):
            for ti, t in enumerate(e.targets):
                p = f"text_model.encoder.layers.{l}.self_attn.{t}."
                yield (p + "lora_A.default.weight", st.A(l)[ti * r:(ti + 1) * r],
                       st.A(l, st.grads)[ti * r:(ti + 1) * r])
                yield p + "lora_B.default.weight", st.B(l)[ti], st.B(l, st.grads)[ti]

    def named_parameters(self):
        """HF / peft key names.  On the device, LoRA tensors are views into the flat trainable buffer with
        ``.grad`` viewing the flat gradient buffer; ``token_embedding.weight`` is the dense [V, D] matrix
        (frozen rows with the lazy decay applied) and its ``.grad`` covers the added rows only — the rows the
        reference leaves non-zero after train_textboost.py:1109-1117."""
        if self._engine is None:
            for k, v in self._sd.items():
                v.requires_grad_(False)
                yield k, v
            return
        live = {}
        for k, v, g in self._trainable_views():
            v.grad = g
            v.requires_grad_(True)
            live[k] = v
        for k, v in self._sd.items():
            if k == "text_model.embeddings.token_embedding.weight":
                w = self.get_input_embeddings().weight
                yield k, w
            else:
                yield k, v
        yield from live.items()

    def parameters(self):
        for _, v in self.named_parameters():
            yield v

    def state_dict(self) -> Dict[str, torch.Tensor]:
        sd = {k: v.detach() for k, v in self.named_parameters()}
        sd["null_embedding"] = self.null_embedding
        return sd

    def added_rows(self) -> torch.Tensor:
        """[n_added, D] view of the trainable embedding rows (ids >= the original vocabulary size)."""
        return self.engine.state.rows()

    # ------------------------------------------------------------------ forward (text_encoder.py:34-87)
    def forward(self, input_ids: Optional[torch.Tensor] = None, attention_mask: Optional[torch.Tensor] = None,
                position_ids: Optional[torch.Tensor] = None, output_attentions: Optional[bool] = None,
                output_hidden_states: Optional[bool] = None, return_dict: Optional[bool] = None):
        if input_ids is None:
            raise ValueError("You have to specify input_ids")
        if attention_mask is not None:
            # the reference never reaches this (utils.py:14-17 drops the mask; with the flag set its mask is a
            # Python list and the call fails): only the causal-mask path exists.
            raise NotImplementedError("attention_mask is not supported on the TextBoost path (causal mask only)")
        if position_ids is not None or output_attentions or output_hidden_states:
            raise NotImplementedError("position_ids / output_attentions / output_hidden_states are not part of "
                                      "the TextBoost path")
        e = self.engine
        ids = input_ids.to(e.device).view(-1, input_ids.shape[-1])
        trainable = torch.is_grad_enabled() and (e.r > 0 or e.state.n_rows > 0) and \
            (self._lora is not None or self._train_embedding)
        if trainable:
            e.pack_lora()
            h = _ClipFunction.apply(self._anchor, e, ids)
        else:
            with torch.no_grad():
                e.pack_lora()
                h = e.forward(ids, save_for_backward=False)
        # pooled output (unused by the reference, utils.py:24 takes [0]): hidden state at the first EOS
        eos = (ids == EOS_ID).int().argmax(dim=-1)
        pooled = h.detach()[torch.arange(ids.shape[0], device=h.device), eos]
        if self._dtype in (torch.float16, torch.bfloat16) and not trainable:
            h = h.to(POLICY.act)
        out = ModelOutput(h, pooled)
        return out if (return_dict if return_dict is not None else True) else tuple(out)

    __call__ = forward


class CLIPTextModel(TextBoostModel):
    """The stock ``transformers.CLIPTextModel`` surface of /root/reference/inference.py:46-58 (pipeline.text_encoder
    + ``load_adapter`` / ``set_adapter``): the same engine without the null-embedding override, inference only."""
    _null_override = False

    def set_null_embedding(self, null_embedding):
        raise AttributeError("CLIPTextModel has no null embedding: use TextBoostModel")


def _save_tensors(sd, directory, stem, safe):
    if safe:
        from safetensors.torch import save_file
        save_file(sd, os.path.join(directory, stem + ".safetensors"), metadata={"format": "pt"})
    else:
        torch.save(sd, os.path.join(directory, stem + ".bin"))


def config_to_dict(cfg: ClipConfig) -> dict:
    return {k: getattr(cfg, k) for k in _CFG_KEYS}


__all__ = ["TextBoostModel", "ModelOutput", "config_to_dict"]
