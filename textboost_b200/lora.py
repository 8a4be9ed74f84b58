"""peft.LoraConfig stand-in for the reference's use of it (/root/reference/train_textboost.py:702-709): rank-r
LoRA on the CLIP attention projections, lora_alpha = r, gaussian init, dropout 0.  Default targets are the
reference's q/k/v; any subset of {q,k,v,out}_proj and any rank 1..16 (``--lora_rank``, run_textboost_db.py:137) is
accepted.  The arithmetic (peft tuners/lora/layer.py Linear.forward, y = W x + b + (alpha/r) B(A(x))) is fused into
the projection GEMMs by textboost_b200.clip.ClipEngine."""
from __future__ import annotations

import dataclasses
import json
import os
from typing import Sequence, Union

DEFAULT_TARGETS = ("q_proj", "k_proj", "v_proj")
SUPPORTED_TARGETS = ("q_proj", "k_proj", "v_proj", "out_proj")


@dataclasses.dataclass
class LoraConfig:
    r: int = 8
    lora_alpha: int = 8
    init_lora_weights: Union[bool, str] = True   # "gaussian": A ~ N(0, (1/r)^2), B = 0
    target_modules: Sequence[str] = DEFAULT_TARGETS
    lora_dropout: float = 0.0
    bias: str = "none"

    def validate(self):
        t = list(self.target_modules)
        if not t or len(set(t)) != len(t) or any(m not in SUPPORTED_TARGETS for m in t):
            raise NotImplementedError(
                f"target_modules={t}: the fused LoRA path covers any subset of {list(SUPPORTED_TARGETS)} "
                "(the reference uses q/k/v, train_textboost.py:705; fc1 / fc2 of its commented-out list are not built)")
        if not 0 < self.r <= 16:
            raise NotImplementedError("LoRA rank must be 1..16 (K extension of the projection GEMMs: <= 64 columns)")
        if self.lora_dropout != 0.0 or self.bias != "none":
            raise NotImplementedError("lora_dropout / bias are not used by the reference path")

    # adapter_config.json as transformers' PeftAdapterMixin.save_pretrained writes it (the keys inference.py's
    # load_adapter needs)
    def to_dict(self):
        return {"peft_type": "LORA", "task_type": None, "base_model_name_or_path": None, "r": self.r,
                "lora_alpha": self.lora_alpha, "lora_dropout": self.lora_dropout, "bias": self.bias,
                "init_lora_weights": self.init_lora_weights, "target_modules": sorted(self.target_modules),
                "fan_in_fan_out": False, "inference_mode": True, "modules_to_save": None}

    def save(self, directory: str):
        with open(os.path.join(directory, "adapter_config.json"), "w") as f:
            json.dump(self.to_dict(), f, indent=2, sort_keys=True)

    @classmethod
    def load(cls, directory: str) -> "LoraConfig":
        with open(os.path.join(directory, "adapter_config.json")) as f:
            d = json.load(f)
        return cls(r=d["r"], lora_alpha=d["lora_alpha"], init_lora_weights=d.get("init_lora_weights", True),
                   target_modules=tuple(d["target_modules"]), lora_dropout=d.get("lora_dropout", 0.0),
                   bias=d.get("bias", "none"))
