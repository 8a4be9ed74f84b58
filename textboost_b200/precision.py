"""Precision policy of the process: which build of the CUDA library is loaded and the torch dtype of every 16-bit
buffer handed to it.

The reference picks ``weight_dtype`` from ``--mixed_precision`` (train_textboost.py:928-933): fp16 -> torch.float16,
bf16 -> torch.bfloat16, and casts the frozen UNet / original text encoder to it (:937-939) while accelerate's autocast
runs the trainable encoder in the same type.  Here the two policies are two builds of the same kernel sources
(libtextboost_b200.so and, compiled with -DTB_BF16, libtextboost_b200_bf16.so: tensor-core operands, saved
activations and frozen weights in the 16-bit type; accumulators, softmax, statistics, master weights and the optimiser
in fp32 in both).  One process uses one policy: set it (``set_policy`` or TEXTBOOST_B200_PRECISION=fp16|bf16) before the
first engine is built; changing it after the library has been loaded raises.

Differences of the bf16 policy that follow the reference: no GradScaler (accelerate only creates one for fp16), so the
loss scale is 1 and no step is ever skipped.  ``--mixed_precision no`` (fp32 weights and activations) is not built.
"""
from __future__ import annotations

import os

import torch

_TABLE = {
    "fp16": (torch.float16, "libtextboost_b200.so", 0),
    "bf16": (torch.bfloat16, "libtextboost_b200_bf16.so", 1),
}


class _Policy:
    def __init__(self):
        name = os.environ.get("TEXTBOOST_B200_PRECISION", "fp16")
        if name not in _TABLE:
            raise ValueError(f"TEXTBOOST_B200_PRECISION={name!r}: expected one of {sorted(_TABLE)}")
        self.name = name
        self.locked = False  # set by _cabi.lib() once a library has been loaded

    @property
    def act(self) -> torch.dtype:
        """torch dtype of the 16-bit tensors (activations, frozen weights, tensor-core operands)."""
        return _TABLE[self.name][0]

    @property
    def lib_name(self) -> str:
        return _TABLE[self.name][1]

    @property
    def storage_code(self) -> int:
        """What tb_storage_dtype() of the matching library returns (TB_STORAGE_F16 / TB_STORAGE_BF16)."""
        return _TABLE[self.name][2]

    @property
    def uses_grad_scaler(self) -> bool:
        return self.name == "fp16"


POLICY = _Policy()


def set_policy(name: str) -> None:
    """Select the policy ('fp16' or 'bf16').  'no' (fp32) raises NotImplementedError."""
    if name == "no":
        raise NotImplementedError("--mixed_precision no (fp32 weights and activations) is not built: the B200 path "
                                  "keeps 16-bit tensor-core operands (fp16 or bf16) with fp32 accumulation")
    if name not in _TABLE:
        raise ValueError(f"precision policy {name!r}: expected one of {sorted(_TABLE)}")
    if POLICY.locked and name != POLICY.name:
        raise RuntimeError(f"the {POLICY.name} library is already loaded in this process; start a new process for "
                           f"{name} (TEXTBOOST_B200_PRECISION={name})")
    POLICY.name = name
