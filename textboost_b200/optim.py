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
LOSS_SCALE, GROWTH_TRACKER, FOUND_INF, SUM_SQ, STEP, FROZEN_DECAY, CLIP_COEF, GRAD_NORM, SKIPPED = range(9)


class FusedAdamW:
    """AdamW over the flat [LoRA | added embedding rows] buffer of a ClipEngine."""

    def __init__(self, engine, lr=5e-5, emb_lr=1e-3, betas=(0.9, 0.999), weight_decay=1e-2, eps=1e-8,
                 max_grad_norm: Optional[float] = 1.0, mean_norm: Optional[float] = None, mixing=None,
                 mixed_precision="fp16", world_size: int = 1):
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
        self.state[LOSS_SCALE] = 65536.0 if mixed_precision == "fp16" else 1.0  # GradScaler init_scale
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
               self.max_grad_norm, 1.0 / self.world_size, self.mean_norm, C.ptr(self.state),
               C.ptr(self.added_norm), C.stream_ptr())

    def zero_grad(self, set_to_none: bool = True):
        """No-op: tb_adamw_fused_step consumes AND zeroes the gradient buffer."""

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
