"""oracle (test infrastructure): one TextBoost training step in plain PyTorch + autograd.

Follows /root/reference/train_textboost.py line by line:
  :1041-1052  noise / timesteps / add_noise           (inputs here, so both samplers are testable)
  :1054-1067  encode_prompt -> unet(...).sample
  :1070-1075  target (epsilon | v_prediction)
  :1077-1094  --with_image_prior: [instance | class] halves, loss = mse(instance) + image_ppl_weight * mse(class)
  :1085-1090  loss = mse(pred.float(), target.float()).mean()
  :1096-1106  knowledge-preservation loss (cos | mse) on the prior prompts, weight kpl_weight
  :1108       backward
  :1109-1117  zero the embedding-gradient rows below min(added_token_ids)
  :1119-1126  --mixing mask on lora_B grads
  :1128-1133  clip_grad_norm_(text_model.encoder.parameters(), max_grad_norm)   (LoRA only)
  :1134-1136  AdamW step (embedding lr = emb_learning_rate, LoRA lr = learning_rate), zero_grad
  :1138-1149  renormalise the added rows to norm <= mean_norm
By default the fp16 autocast / GradScaler of accelerate is NOT modelled: the oracle is the exact-arithmetic
(fp32, or fp64 when the caller casts the modules) statement of the same function.  With
``mixed_precision="fp16"`` (GPU only) the same statements run under the reference's precision policy
(:928-939, :1063-1066): trainable text encoder fp32 under autocast(fp16) with fp32
outputs, UNet and frozen text encoder cast to fp16 by the caller and run without autocast, fp32 loss, scaled
backward + unscale -- i.e. what torch's own fp16 kernels make of the reference path.  Tests use it to report the
north star's rtol 1e-3 / atol 1e-4 criterion against "the reference fp16 path" as well as against fp32.
"""
from __future__ import annotations

import contextlib
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import ddpm_ref


def lora_named_parameters(te):
    return [(n, p) for n, p in te.named_parameters() if "lora_" in n]


def forward_loss(unet, te, te0, latents, noise, timesteps, input_ids, prior_ids=None, kpl_weight=0.1,
                 kpl_type="cos", prediction_type="epsilon", image_ppl_weight=None, mixed_precision=None):
    """Returns (loss, model_pred, encoder_hidden_states).  image_ppl_weight (not None = --with_image_prior,
    :1077-1094): the batch is [instance | class] halves, loss = mse(instance) + image_ppl_weight * mse(class)."""
    assert mixed_precision in (None, "fp16", "bf16")
    wdt = {None: None, "fp16": torch.float16, "bf16": torch.bfloat16}[mixed_precision]  # weight_dtype, :928-933
    amp = (lambda: torch.autocast(latents.device.type, dtype=wdt)) if mixed_precision else \
        contextlib.nullcontext
    noisy = ddpm_ref.add_noise(latents, noise, timesteps)
    with amp():
        ehs = te(input_ids)
    ehs = ehs.float()  # accelerate's convert_outputs_to_fp32
    if mixed_precision:
        pred = unet(noisy.to(wdt), timesteps, ehs.to(wdt))  # :1063-1066 (weight_dtype casts)
    else:
        pred = unet(noisy, timesteps, ehs)
    if prediction_type == "epsilon":
        target = noise
    elif prediction_type == "v_prediction":
        target = ddpm_ref.get_velocity(latents, noise, timesteps)
    else:
        raise ValueError(prediction_type)
    if image_ppl_weight is not None:
        pred_i, pred_c = torch.chunk(pred, 2, dim=0)
        target_i, target_c = torch.chunk(target, 2, dim=0)
        prior_loss = F.mse_loss(pred_c.float(), target_c.float(), reduction="mean")
        loss = F.mse_loss(pred_i.float(), target_i.float(), reduction="none").mean() + image_ppl_weight * prior_loss
    else:
        loss = F.mse_loss(pred.float(), target.float(), reduction="none").mean()
    if kpl_weight > 0.0 and prior_ids is not None:
        with amp():
            h = te(prior_ids)
        h = h.float()
        with torch.no_grad():
            h0 = te0(prior_ids).float()
        if kpl_type == "cos":
            kp = (1 - F.cosine_similarity(h, h0, dim=-1)).mean()
        else:
            kp = F.mse_loss(h, h0, reduction="mean")
        loss = loss + kpl_weight * kp
    return loss, pred, ehs


def reference_step(unet, te, te0, latents, noise, timesteps, input_ids, prior_ids=None, *,
                   n_base: int, kpl_weight=0.1, kpl_type="cos", prediction_type="epsilon",
                   optimizer: Optional[torch.optim.Optimizer] = None, max_grad_norm=1.0, mixing=None,
                   mean_norm: Optional[float] = None, image_ppl_weight=None, mixed_precision=None,
                   loss_scale: float = 1.0) -> Dict[str, torch.Tensor]:
    """Runs forward + backward (+ the optimiser tail when `optimizer` is given).

    n_base = min(added_token_ids): rows below it never train.  Returns loss, pred, d(ehs) and the
    gradients that reach the trainable parameters (after the row mask, before clipping).
    """
    emb = te.get_input_embeddings().weight
    for _, p in lora_named_parameters(te):
        p.grad = None
    emb.grad = None
    unet_lora = [(n, p) for n, p in unet.named_parameters() if p.requires_grad]  # crossattn_kv adapter (:712-721)
    for _, p in unet_lora:
        p.grad = None
    loss, pred, ehs = forward_loss(unet, te, te0, latents, noise, timesteps, input_ids, prior_ids,
                                   kpl_weight, kpl_type, prediction_type, image_ppl_weight, mixed_precision)
    ehs.retain_grad()
    (loss * loss_scale).backward()  # GradScaler.scale(loss).backward(); unscale_ precedes clipping (:1128-1133)
    if loss_scale != 1.0:
        inv = 1.0 / loss_scale
        ehs.grad.mul_(inv)
        if emb.grad is not None:
            emb.grad.mul_(inv)
        for _, p in lora_named_parameters(te):
            if p.grad is not None:
                p.grad.mul_(inv)
        for _, p in unet_lora:
            if p.grad is not None:
                p.grad.mul_(inv)
    if emb.grad is not None:
        emb.grad[:n_base] = 0  # :1109-1117
    if mixing is not None:  # :1119-1126
        for n, p in lora_named_parameters(te):
            if "lora_B" in n:
                if mixing == "object":
                    p.grad[1::2, :] = 0.0
                else:
                    p.grad[0::2, :] = 0.0
    out = {"loss": loss.detach(), "pred": pred.detach(), "d_ehs": ehs.grad.detach().clone(),
           "grad_rows": emb.grad[n_base:].detach().clone() if emb.grad is not None else None,
           "grad_lora": {n: p.grad.detach().clone() for n, p in lora_named_parameters(te)},
           "grad_unet_lora": {n: p.grad.detach().clone() for n, p in unet_lora if p.grad is not None}}
    if optimizer is not None:
        if max_grad_norm:
            out["grad_norm"] = torch.nn.utils.clip_grad_norm_(
                [p for _, p in lora_named_parameters(te)], max_grad_norm)
        optimizer.step()
        optimizer.zero_grad(set_to_none=True)
        if mean_norm is not None:  # :1138-1149
            with torch.no_grad():
                rows = emb[n_base:]
                v = rows.norm(dim=-1, keepdim=True)
                out["added_embedding_norm"] = v.mean()
                emb[n_base:] = torch.minimum(torch.full_like(v, mean_norm), v) / v * rows
    return out


def make_optimizer(te, learning_rate=5e-5, emb_learning_rate=1e-3, betas=(0.9, 0.999), weight_decay=1e-2,
                   eps=1e-8, unet=None):
    """train_textboost.py:829-854; `unet` with trainable (LoRA) parameters adds the third group of :838-841."""
    groups = [{"params": [te.get_input_embeddings().weight], "lr": emb_learning_rate},
              {"params": [p for _, p in lora_named_parameters(te)]}]
    if unet is not None and any(p.requires_grad for p in unet.parameters()):
        groups.append({"params": [p for p in unet.parameters() if p.requires_grad]})
    return torch.optim.AdamW(
        groups,
        lr=learning_rate, betas=betas, weight_decay=weight_decay, eps=eps)
