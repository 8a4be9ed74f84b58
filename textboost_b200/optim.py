"""Fused optimiser tail of the TextBoost step behind a torch.optim-like surface.

One C-ABI call (tb_adamw_fused_step, three small kernels) replaces, with identical results,
/root/reference/train_textboost.py:1109-1149 plus what accelerate does around it:

  GradScaler.unscale_ + inf check  ->  zero embedding-grad rows below min(added_token_ids) (:1109-1117; those
  rows never exist here)  ->  --mixing mask on lora_B grads (:1119-1126)  ->  clip_grad_norm_ over the LoRA
  parameters only (:1128-1133)  ->  torch.optim.AdamW with two learning rates, decoupled weight decay that
  also shrinks every frozen embedding row (:829-854, SURVEY.md D8: kept as one device scalar)  ->
  zero_grad (:1136)  ->  renormalise the added rows to norm <= mean_norm (:1138-1149)  ->  GradScaler.update.

All control scalars live in a device fp32[16] vector, so step() never synchronises and can sit inside a
captured CUDA graph.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _cabi as C

F32 = torch.float32

# indices into the device state vector (csrc/optim.cu)
LOSS_SCALE, GROWTH_TRACKER, FOUND_INF, SUM_SQ, STEP, FROZEN_DECAY, CLIP_COEF, GRAD_NORM, SKIPPED, LR_MULT, FIXED_SCALE = range(11)

# --lr_scheduler names of the reference CLI (train_textboost.py:224-231 -> diffusers.optimization.get_scheduler)
LR_SCHEDULES = {"constant": 0, "constant_with_warmup": 1, "linear": 2, "cosine": 3, "cosine_with_restarts": 4,
                "polynomial": 5}


def lr_multiplier(name: str, step: int, warmup: int, total: int, lr_init: float = 1.0) -> float:
    """Host statement of the schedule the optimiser kernel evaluates on the device (csrc/optim.cu lr_multiplier):
    diffusers.optimization's LambdaLR factories with their defaults.  `step` = successful optimiser steps so far."""
    import math
    kind = LR_SCHEDULES[name]
    if kind == 0:
        return 1.0
    if step < warmup:
        return step / max(1.0, warmup)
    if kind == 1:
        return 1.0
    progress = (step - warmup) / max(1.0, total - warmup)
    if kind == 2:
        return max(0.0, (total - step) / max(1.0, total - warmup))
    if kind == 3:
        return max(0.0, 0.5 * (1.0 + math.cos(math.pi * progress)))
    if kind == 4:
        return 0.0 if progress >= 1.0 else max(0.0, 0.5 * (1.0 + math.cos(math.pi * (progress % 1.0))))
    lr_end = 1e-7
    if step > total:
        return lr_end / lr_init
    return ((lr_init - lr_end) * (1.0 - (step - warmup) / (total - warmup)) + lr_end) / lr_init


class FusedAdamW:
    """AdamW over the flat [LoRA | added embedding rows] buffer of a ClipEngine."""

    def __init__(self, engine, lr=5e-5, emb_lr=1e-3, betas=(0.9, 0.999), weight_decay=1e-2, eps=1e-8,
                 max_grad_norm: Optional[float] = 1.0, mean_norm: Optional[float] = None, mixing=None,
                 mixed_precision="fp16", world_size: int = 1, lr_scheduler: str = "constant",
                 lr_warmup_steps: int = 0, max_train_steps: int = 0, gradient_accumulation_steps: int = 1):
        if lr_scheduler not in LR_SCHEDULES:
            raise ValueError(f"unknown lr_scheduler {lr_scheduler!r}; one of {sorted(LR_SCHEDULES)}")
        self.lr_scheduler, self.lr_warmup_steps, self.max_train_steps = lr_scheduler, lr_warmup_steps, max_train_steps
        # accelerate divides the loss by gradient_accumulation_steps before backward; the micro-batch gradients are
        # summed in the flat buffer here, so the division joins the 1/world of the all-reduce average
        self.gradient_accumulation_steps = int(gradient_accumulation_steps)
        self.engine = engine
        st = engine.state
        self.betas, self.weight_decay, self.eps = tuple(betas), weight_decay, eps
        self.max_grad_norm = max_grad_norm if max_grad_norm is not None else 0.0
        self.mixing = mixing
        self.world_size = world_size
        # same group order as train_textboost.py:829-837: embedding first, then the encoder (LoRA) params
        self.param_groups = [{"name": "token_embedding", "lr": emb_lr, "params": [st.rows()]},
                             {"name": "lora", "lr": lr, "params": [st.params[:st.n_lora]]}]
        self.exp_avg = torch.zeros_like(st.params)
        self.exp_avg_sq = torch.zeros_like(st.params)
        self.state = torch.zeros(16, device=st.params.device, dtype=F32)
        # accelerate creates a GradScaler for fp16 only (init_scale 65536, growth 2x / 2000 steps, backoff 0.5); under
        # bf16 the loss is not scaled and the scale never moves
        self.state[LOSS_SCALE] = 65536.0 if mixed_precision == "fp16" else 1.0
        self.state[FIXED_SCALE] = 0.0 if mixed_precision == "fp16" else 1.0
        self.state[FROZEN_DECAY] = 1.0
        engine.decay = self.state[FROZEN_DECAY:FROZEN_DECAY + 1]
        if mean_norm is None:
            # train_textboost.py:1017: mean row norm of the resized embedding matrix
            n = engine.tok_base.norm(dim=-1).sum() + st.rows().norm(dim=-1).sum()
            mean_norm = float(n / (engine.tok_base.shape[0] + st.n_rows))
        self.mean_norm = mean_norm
        self.added_norm = torch.zeros(1, device=st.params.device, dtype=F32)

    # GradScaler surface ------------------------------------------------------------------------------
    @property
    def loss_scale(self) -> torch.Tensor:
        """Device scalar (fp32[1]) the loss gradient must be multiplied by (accelerator.backward)."""
        return self.state[LOSS_SCALE:LOSS_SCALE + 1]

    def scale(self, loss: torch.Tensor) -> torch.Tensor:
        return loss * self.loss_scale

    # torch.optim surface -----------------------------------------------------------------------------
    def step(self):
        st = self.engine.state
        if self.mixing is not None and st.n_b:
            parity = 1 if self.mixing == "object" else 0  # train_textboost.py:1119-1126
            C.call("tb_optim_mix_mask", C.ptr(st.b_segment(st.grads)), st.n_b, st.D, st.r, parity, C.stream_ptr())
        C.call("tb_adamw_fused_step", C.ptr(st.params), C.ptr(st.grads), C.ptr(self.exp_avg),
               C.ptr(self.exp_avg_sq), st.n_lora, st.n_rows, st.D, float(self.param_groups[1]["lr"]),
               float(self.param_groups[0]["lr"]), self.betas[0], self.betas[1], self.eps, self.weight_decay,
               self.max_grad_norm, 1.0 / (self.world_size * self.gradient_accumulation_steps), self.mean_norm,
               LR_SCHEDULES[self.lr_scheduler], float(self.lr_warmup_steps), float(self.max_train_steps),
               C.ptr(self.state), C.ptr(self.added_norm), C.stream_ptr())

    def get_last_lr(self):
        """[embedding lr, LoRA lr] of the step just taken (lr_scheduler.get_last_lr of the reference, :1230); reads
        the device state vector, i.e. synchronises."""
        m = float(self.state[LR_MULT]) if int(self.state[STEP]) > 0 else lr_multiplier(
            self.lr_scheduler, 0, self.lr_warmup_steps, self.max_train_steps)
        return [g["lr"] * m for g in self.param_groups]

    def zero_grad(self, set_to_none: bool = True):
        """No-op: tb_adamw_fused_step consumes AND zeroes the gradient buffer."""

    def hyperparameters(self):
        """Everything step() passes to the kernel by value (what a captured CUDA graph bakes in)."""
        return (float(self.param_groups[1]["lr"]), float(self.param_groups[0]["lr"]), tuple(self.betas), self.eps,
                self.weight_decay, self.max_grad_norm, self.mean_norm, self.lr_scheduler, self.lr_warmup_steps,
                self.max_train_steps, self.gradient_accumulation_steps, self.world_size)

    def state_dict(self):
        return {"exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                "state": self.state.clone(), "mean_norm": self.mean_norm,
                "lr": self.param_groups[1]["lr"], "emb_lr": self.param_groups[0]["lr"]}

    def load_state_dict(self, sd):
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.state.copy_(sd["state"])
        self.mean_norm = sd["mean_norm"]
        self.param_groups[1]["lr"], self.param_groups[0]["lr"] = sd["lr"], sd["emb_lr"]


class FlatAdamW:
    """AdamW over one flat fp32 buffer with one learning rate, no gradient clipping and no embedding rows: the third
    parameter group of train_textboost.py:838-841 (the UNet LoRA of --unet_params_to_train crossattn_kv; clipping
    covers the text encoder only, :1128-1133).  The same tb_adamw_fused_step, with an empty row segment and
    max_grad_norm 0.  No GradScaler of its own (the loss scale of the step is 1: the mode runs under the bf16 policy,
    see TextBoostTrainer)."""

    def __init__(self, params: torch.Tensor, grads: torch.Tensor, lr=5e-5, betas=(0.9, 0.999), weight_decay=1e-2,
                 eps=1e-8, world_size: int = 1, lr_scheduler: str = "constant", lr_warmup_steps: int = 0,
                 max_train_steps: int = 0, gradient_accumulation_steps: int = 1):
        assert params.dtype == F32 and grads.shape == params.shape and params.is_contiguous()
        self.params, self.grads = params, grads
        self.lr, self.betas, self.weight_decay, self.eps = lr, tuple(betas), weight_decay, eps
        self.world_size, self.gradient_accumulation_steps = world_size, int(gradient_accumulation_steps)
        self.lr_scheduler, self.lr_warmup_steps, self.max_train_steps = lr_scheduler, lr_warmup_steps, max_train_steps
        self.exp_avg = torch.zeros_like(params)
        self.exp_avg_sq = torch.zeros_like(params)
        self.state = torch.zeros(16, device=params.device, dtype=F32)
        self.state[LOSS_SCALE] = 1.0
        self.state[FIXED_SCALE] = 1.0
        self.state[FROZEN_DECAY] = 1.0
        self._norm_out = torch.zeros(1, device=params.device, dtype=F32)

    def step(self):
        C.call("tb_adamw_fused_step", C.ptr(self.params), C.ptr(self.grads), C.ptr(self.exp_avg),
               C.ptr(self.exp_avg_sq), self.params.numel(), 0, 0, float(self.lr), float(self.lr), self.betas[0],
               self.betas[1], self.eps, self.weight_decay, 0.0,
               1.0 / (self.world_size * self.gradient_accumulation_steps), 0.0, LR_SCHEDULES[self.lr_scheduler],
               float(self.lr_warmup_steps), float(self.max_train_steps), C.ptr(self.state), C.ptr(self._norm_out),
               C.stream_ptr())

    def hyperparameters(self):
        return (float(self.lr), self.betas, self.eps, self.weight_decay, self.lr_scheduler, self.lr_warmup_steps,
                self.max_train_steps, self.gradient_accumulation_steps, self.world_size)

    def state_dict(self):
        return {"exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(), "state": self.state.clone(),
                "lr": self.lr}

    def load_state_dict(self, sd):
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.state.copy_(sd["state"])
        self.lr = sd["lr"]
