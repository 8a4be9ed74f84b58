"""oracle (test infrastructure): build the oracle twin of a synthetic TextBoost trainer and compare one step.

Used only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs: it is
the checker (or the timed CPU baseline), never something the product path routes through.

``twin_of(trainer)`` rebuilds, from the state dicts a ``synthetic.build_trainer(..., keep_sd=True)`` kept,
the plain-PyTorch modules of oracle/{unet,clip}_ref.py with identical weights, LoRA factors, added
embedding rows and null embedding; ``compare_step`` runs oracle/step_ref.reference_step
(train_textboost.py:1041-1149) and the product's step on the same batch and returns the error of every
quantity the north star names: noise prediction, loss, LoRA A/B gradients, added-row gradients, and the
parameters after the optimiser tail.
"""
from __future__ import annotations

from typing import Dict

import torch

from . import clip_ref, step_ref, unet_ref

LORA_TARGETS = ("q_proj", "k_proj", "v_proj")  # the reference's targets (train_textboost.py:705)


def targets_of(trainer):
    """LoRA targets of the product engine, in the order of its flat trainable buffer."""
    return tuple(getattr(trainer.te, "targets", LORA_TARGETS)) or LORA_TARGETS


def _unet_cfg(c) -> unet_ref.UNetConfig:
    return unet_ref.UNetConfig(
        in_channels=c.in_channels, out_channels=c.out_channels, block_out_channels=tuple(c.block_out_channels),
        layers_per_block=c.layers_per_block, cross_attention_dim=c.cross_attention_dim,
        attention_head_dim=tuple(c.attention_head_dim), down_has_attn=tuple(c.down_has_attn),
        norm_num_groups=c.norm_num_groups, norm_eps=c.norm_eps, use_linear_projection=c.use_linear_projection,
        sample_size=c.sample_size)


def _clip_cfg(c, vocab=None) -> clip_ref.ClipTextConfig:
    return clip_ref.ClipTextConfig(
        vocab_size=vocab or c.vocab_size, hidden_size=c.hidden_size, intermediate_size=c.intermediate_size,
        num_hidden_layers=c.num_hidden_layers, num_attention_heads=c.num_attention_heads,
        hidden_act=c.hidden_act)


def rel_max(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b|: the scale-free error used for fp16-vs-fp32 comparisons."""
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def pass_fraction(a: torch.Tensor, b: torch.Tensor, rtol=1e-3, atol=1e-4) -> float:
    """fraction of elements with |a - b| <= atol + rtol |b| (torch.allclose's criterion, the north star's
    rtol 1e-3 / atol 1e-4), reported rather than asserted as all-or-nothing."""
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).abs() <= atol + rtol * b.abs()).float().mean().item()


def twin_of(trainer, device="cpu", dtype=torch.float32, frozen_dtype=None):
    """(unet, te, te0, optimizer) oracle modules with the trainer's current weights.  frozen_dtype (fp16): the
    reference's weight_dtype cast of the UNet and the frozen text encoder (train_textboost.py:937-939)."""
    frozen_dtype = dtype if frozen_dtype is None else frozen_dtype
    syn = trainer.synthetic
    assert "unet_sd" in syn, "build the trainer with keep_sd=True"
    ucfg, ccfg = syn["unet_cfg"], syn["clip_cfg"]
    V, n_added, r = ccfg.vocab_size, syn["n_added"], syn["lora_r"]
    unet = unet_ref.UNet2DConditionModelRef(_unet_cfg(ucfg))
    unet.load_state_dict({k: v.detach().to("cpu", torch.float32) for k, v in syn["unet_sd"].items()})
    unet = unet.to(device, frozen_dtype).requires_grad_(False)
    kvl = getattr(trainer.unet, "kv_lora", None)
    if kvl is not None:  # --unet_params_to_train crossattn_kv: same adapter weights on the oracle UNet
        mods = unet.add_cross_kv_lora(kvl.r, kvl.scaling * kvl.r)
        assert [n for n, _ in mods] == [n for n in kvl.names], "adapter order differs from the engine's"
        sd = kvl.state_dict()
        with torch.no_grad():
            for n, m in mods:
                m.lora_A["default"].weight.copy_(sd[n + ".lora_A.weight"])
                m.lora_B["default"].weight.copy_(sd[n + ".lora_B.weight"])
        unet = unet.to(device, frozen_dtype)  # the reference casts the adapter with the UNet (:937)
    csd = {k: v.detach().to("cpu", torch.float32) for k, v in syn["clip_sd"].items()}
    te_eng = trainer.te
    null = te_eng.null_embedding.detach().to("cpu", torch.float32)
    te0 = None
    if trainer.te0 is not None:
        te0 = clip_ref.TextBoostModelRef(_clip_cfg(ccfg))
        te0.load_state_dict(csd, strict=False)
        te0.set_null_embedding(null.clone())
        te0 = te0.to(device, frozen_dtype).requires_grad_(False)
    te = clip_ref.TextBoostModelRef(_clip_cfg(ccfg))
    te.load_state_dict(csd, strict=False)
    te.resize_token_embeddings(V + n_added)
    st = te_eng.state
    with torch.no_grad():
        te.get_input_embeddings().weight[V:] = st.rows().detach().cpu()
    te.set_null_embedding(null.clone())
    te.requires_grad_(False)
    if r:
        te.add_adapter(r=r, lora_alpha=getattr(te_eng, "scaling", 1.0) * r, target_modules=targets_of(trainer))
        with torch.no_grad():
            for l, lyr in enumerate(te.text_model.encoder.layers):
                for ti, t in enumerate(targets_of(trainer)):
                    m = getattr(lyr.self_attn, t)
                    m.lora_A["default"].weight.copy_(st.A(l)[ti * r:(ti + 1) * r].cpu())
                    m.lora_B["default"].weight.copy_(st.B(l)[ti].cpu())
    te = te.to(device, dtype)
    te.get_input_embeddings().weight.requires_grad_(True)
    opt = step_ref.make_optimizer(te, learning_rate=trainer.lr, emb_learning_rate=trainer.emb_lr,
                                  betas=(trainer.b1, trainer.b2), weight_decay=trainer.wd, eps=trainer.eps, unet=unet)
    return unet, te, te0, opt


def lora_grads_flat(trainer, ref_grads: Dict[str, torch.Tensor], buf: torch.Tensor):
    """(ours, ref) concatenated LoRA gradient vectors in the same (layer, target, A|B) order, plus the worst
    per-tensor max-relative error."""
    st = trainer.te.state
    r = st.r
    ours, refs, worst = [], [], 0.0
    if not r:  # --lora_rank 0: only the embedding rows train (train_textboost.py:700, 722)
        return torch.zeros(0), torch.zeros(0), 0.0
    for l in range(st.n_layers):
        for ti, t in enumerate(targets_of(trainer)):
            n = f"text_model.encoder.layers.{l}.self_attn.{t}."
            ra = ref_grads[n + "lora_A.default.weight"]
            rb = ref_grads[n + "lora_B.default.weight"]
            oa, ob = st.A(l, buf)[ti * r:(ti + 1) * r], st.B(l, buf)[ti]
            worst = max(worst, rel_max(oa, ra), rel_max(ob, rb))
            ours += [oa.detach().float().cpu().flatten(), ob.detach().float().cpu().flatten()]
            refs += [ra.detach().float().cpu().flatten(), rb.detach().float().cpu().flatten()]
    return torch.cat(ours), torch.cat(refs), worst


def _reference_fp16_policy(trainer, b, V, kind, use_kpl, device, policy="fp16"):
    """The same step through torch's own 16-bit kernels under the reference's precision policy (step_ref
    mixed_precision="fp16" or "bf16"; GradScaler only for fp16): returns (pred, LoRA grad vector, added-row grads)
    on the CPU."""
    wdt = {"fp16": torch.float16, "bf16": torch.bfloat16}[policy]
    unet, te, te0, _ = twin_of(trainer, device, torch.float32, frozen_dtype=wdt)
    ref = step_ref.reference_step(
        unet, te, te0, b["latents"].float(), b["noise"].float(), b["timesteps"], b["input_ids"],
        b["prior_ids"] if use_kpl else None, n_base=V, kpl_weight=trainer.kpl_weight, kpl_type=kind,
        prediction_type="v_prediction" if trainer.v_pred else "epsilon", optimizer=None, mixing=trainer.mixing,
        image_ppl_weight=getattr(trainer, "image_prior_weight", None), mixed_precision=policy,
        loss_scale=65536.0 if policy == "fp16" else 1.0)
    _, gr, _ = lora_grads_flat(trainer, ref["grad_lora"], trainer.te.state.grads)
    rows = ref["grad_rows"].detach().float().cpu() if ref["grad_rows"] is not None else None
    return ref["pred"].detach().float().cpu(), gr, rows


def compare_step(trainer, batch: Dict[str, torch.Tensor], device="cpu", dtype=torch.float32,
                 with_optimizer=True, fp16_reference=False, policy="fp16") -> Dict[str, float]:
    """Run one step on the product (trainer, CUDA) and on the oracle twin; return error metrics.

    fp16_reference (device must be CUDA): also run the oracle under the reference's 16-bit policy (`policy`: "fp16",
    or "bf16" when the product runs its bf16 build) and report the
    north star's elementwise criterion (rtol 1e-3 / atol 1e-4) three ways -- ours vs fp32 oracle, torch-fp16 vs
    fp32 oracle, ours vs torch-fp16 -- as pass fractions (keys tol_*)."""
    ref16 = None
    if fp16_reference:
        V0 = trainer.synthetic["clip_cfg"].vocab_size
        use_kpl0 = trainer.kpl_weight > 0 and trainer.te0 is not None
        b0 = {k: v.to(device) for k, v in batch.items()}
        ref16 = _reference_fp16_policy(trainer, b0, V0, {0: "cos", 1: "mse"}[trainer.kpl_kind], use_kpl0, device,
                                       policy)
        del b0
        if torch.cuda.is_available():
            torch.cuda.empty_cache()
    unet, te, te0, opt = twin_of(trainer, device, dtype)
    V = trainer.synthetic["clip_cfg"].vocab_size
    kind = {0: "cos", 1: "mse"}[trainer.kpl_kind]
    use_kpl = trainer.kpl_weight > 0 and te0 is not None
    b = {k: v.to(device) for k, v in batch.items()}
    ref = step_ref.reference_step(
        unet, te, te0, b["latents"].to(dtype), b["noise"].to(dtype), b["timesteps"], b["input_ids"],
        b["prior_ids"] if use_kpl else None, n_base=V, kpl_weight=trainer.kpl_weight, kpl_type=kind,
        prediction_type="v_prediction" if trainer.v_pred else "epsilon",
        optimizer=opt if with_optimizer else None, max_grad_norm=trainer.max_grad_norm,
        mixing=trainer.mixing, mean_norm=trainer.mean_norm,
        image_ppl_weight=getattr(trainer, "image_prior_weight", None))
    scale = trainer.opt_state[0].item()
    loss = trainer.forward_backward(batch["latents"], batch["noise"], batch["timesteps"], batch["input_ids"],
                                    batch["prior_ids"] if use_kpl else None)
    st = trainer.te.state
    g = st.grads.detach().clone() / scale
    if trainer.mixing is not None:  # the product applies the mask inside optimizer_step
        parity = 1 if trainer.mixing == "object" else 0
        gb = st.b_segment(g).view(-1, st.D, st.r)
        gb[:, parity::2, :] = 0
    out = {"loss": loss.item(), "loss_ref": ref["loss"].item(),
           "pred_rel": rel_max(trainer._pred, ref["pred"])}
    go, gr, worst = lora_grads_flat(trainer, ref["grad_lora"], g)
    if go.numel():
        out["lora_grad_rel_l2"] = ((go - gr).norm() / gr.norm()).item()
        out["lora_grad_cos"] = torch.nn.functional.cosine_similarity(go, gr, dim=0).item()
    else:
        out["lora_grad_rel_l2"], out["lora_grad_cos"] = 0.0, 1.0
    out["lora_grad_worst_tensor_rel"] = worst
    out["lora_grad_ours"], out["lora_grad_ref"] = go, gr
    out["pred_ref"] = ref["pred"].detach().float().cpu()
    if ref["grad_rows"] is not None and st.n_rows:
        out["row_grad_rel"] = rel_max(st.rows(g), ref["grad_rows"])
        out["row_grad_ours"], out["row_grad_ref"] = st.rows(g).detach().cpu(), ref["grad_rows"].detach().cpu()
    kvl = getattr(trainer.unet, "kv_lora", None)
    if kvl is not None:
        gu = kvl.grads.detach().clone() / scale
        ours_u, ref_u = [], []
        off = kvl.off.tolist()
        for i, n in enumerate(kvl.names):
            ours_u += [kvl.A(gu)[i * kvl.r:(i + 1) * kvl.r].flatten().cpu(), kvl.B(gu)[off[i]:off[i + 1]].flatten().cpu()]
            ref_u += [ref["grad_unet_lora"][n + ".lora_A.default.weight"].float().flatten().cpu(),
                      ref["grad_unet_lora"][n + ".lora_B.default.weight"].float().flatten().cpu()]
        uo, ur = torch.cat(ours_u), torch.cat(ref_u)
        out["unet_lora_grad_rel_l2"] = ((uo - ur).norm() / ur.norm()).item()
        out["unet_lora_grad_cos"] = torch.nn.functional.cosine_similarity(uo, ur, dim=0).item()
        out["unet_lora_grad_norm_ref"] = ur.norm().item()
    if ref16 is not None:
        p16, g16, r16 = ref16
        pairs = {"pred": (trainer._pred, ref["pred"], p16), "lora_grad": (go, gr, g16)}
        if r16 is not None and st.n_rows:
            pairs["row_grad"] = (st.rows(g), ref["grad_rows"], r16)
        for name, (ours, r32, r16v) in pairs.items():
            out[f"tol_{name}_ours_vs_fp32"] = pass_fraction(ours, r32)
            out[f"tol_{name}_torch16_vs_fp32"] = pass_fraction(r16v, r32)
            out[f"tol_{name}_ours_vs_torch16"] = pass_fraction(ours, r16v)
            out[f"rel_{name}_torch16_vs_fp32"] = rel_max(r16v, r32)
    if with_optimizer:
        trainer.all_reduce()
        trainer.optimizer_step()
        torch.cuda.synchronize()
        out["grad_norm"], out["grad_norm_ref"] = trainer.opt_state[7].item(), ref["grad_norm"].item()
        out["added_norm"], out["added_norm_ref"] = trainer.added_norm.item(), ref["added_embedding_norm"].item()
        emb = te.get_input_embeddings().weight.detach()
        out["rows_after_rel"] = rel_max(st.rows(), emb[V:])
        p_ours, p_ref = [], []
        r = st.r
        for l, lyr in enumerate(te.text_model.encoder.layers):
            for ti, t in enumerate(targets_of(trainer) if r else ()):
                m = getattr(lyr.self_attn, t)
                p_ours += [st.A(l)[ti * r:(ti + 1) * r].detach().cpu().flatten(), st.B(l)[ti].detach().cpu().flatten()]
                p_ref += [m.lora_A["default"].weight.detach().cpu().flatten(),
                          m.lora_B["default"].weight.detach().cpu().flatten()]
        if p_ours:
            po, pr = torch.cat(p_ours), torch.cat(p_ref)
            out["lora_param_max_abs_diff"] = (po - pr).abs().max().item()
        else:
            out["lora_param_max_abs_diff"] = 0.0
        if kvl is not None:
            pu, pr = [], []
            mods = dict(unet.cross_kv_lora_modules())
            off = kvl.off.tolist()
            for i, n in enumerate(kvl.names):
                pu += [kvl.A()[i * kvl.r:(i + 1) * kvl.r].detach().flatten().cpu(), kvl.B()[off[i]:off[i + 1]].detach().flatten().cpu()]
                pr += [mods[n].lora_A["default"].weight.detach().float().flatten().cpu(),
                       mods[n].lora_B["default"].weight.detach().float().flatten().cpu()]
            out["unet_lora_param_max_abs_diff"] = (torch.cat(pu) - torch.cat(pr)).abs().max().item()
        out["frozen_decay"] = trainer.opt_state[5].item()
        base0 = trainer.synthetic["clip_sd"]["text_model.embeddings.token_embedding.weight"][5].float().cpu()
        out["frozen_decay_ref"] = (emb[5].cpu() / base0).mean().item()
    return out
