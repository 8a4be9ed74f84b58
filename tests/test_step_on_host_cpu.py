"""The whole TextBoost training step on the CPU, through the PRODUCT engines: TextBoostTrainer / UNetEngine / ClipEngine /
FusedAdamW run unchanged on CPU tensors — every SIMT entry point executes the product's CUDA source via the host build of
the C ABI (tests/kernel_host_emulation.py), the four tensor-core wrappers are plain-PyTorch statements of their contracts
(tests/engine_standin.py) — and is compared with the oracle step (oracle/step_ref.py) by the same harness and tolerances
the GPU test uses (tests/test_gpu_step.py::test_step_tiny_vs_oracle).  What this pins on a machine without a GPU: the
hand-derived activation-backward chain of the UNet, the text-encoder forward / backward with LoRA packed into the fused
QKV GEMM, the loss / KPL / optimiser tail, and the --with_image_prior two-part loss (whose GPU tests have not run yet)."""
import pytest
import torch

import engine_standin

# a stuck barrier in the block emulator must fail the test, not hang the suite
pytestmark = pytest.mark.timeout(900, method="thread")


def _compare(monkeypatch, B, lora_r=4, **trainer_kw):
    from oracle import harness
    from textboost_b200 import synthetic
    engine_standin.install(monkeypatch)
    tr = synthetic.build_trainer("tiny", "cpu", seed=1, n_added=2, lora_b_std=0.02 if lora_r else 0.0, keep_sd=True,
                                 learning_rate=1e-4, lora_r=lora_r, **trainer_kw)
    V = tr.synthetic["clip_cfg"].vocab_size
    bt = synthetic.batch(B, 8, 3, V, "cpu")
    bt["input_ids"][1, 4] = V + 1
    bt["prior_ids"][0, 1:] = synthetic.EOS  # an empty prior prompt: the null-embedding override path
    return harness.compare_step(tr, bt), tr, bt


def _check(r, tr):
    assert abs(r["loss"] - r["loss_ref"]) < 2e-3 * abs(r["loss_ref"])
    assert r["pred_rel"] < 4e-3
    assert r["lora_grad_rel_l2"] < 5e-3 and r["lora_grad_cos"] > 0.99999
    assert r["row_grad_rel"] < 5e-3
    assert abs(r["grad_norm"] - r["grad_norm_ref"]) < 3e-3 * r["grad_norm_ref"]
    # Adam's first step is lr * sign(g): an element whose gradient is ~0 may step the other way, and the order of the
    # emulated atomics varies from run to run, so the row norms agree to a few 1e-4 only
    assert abs(r["added_norm"] - r["added_norm_ref"]) < 2e-3 * r["added_norm_ref"]
    assert abs(r["frozen_decay"] - r["frozen_decay_ref"]) < 1e-6
    assert r["lora_param_max_abs_diff"] <= 2.1 * tr.lr


def test_whole_step_on_cpu_matches_oracle(monkeypatch):
    r, tr, _ = _compare(monkeypatch, 2, kpl_type="cos")
    _check(r, tr)


@pytest.mark.long_cpu
def test_image_prior_step_on_cpu_matches_oracle(monkeypatch):
    """--with_image_prior (train_textboost.py:1077-1094): [instance | class] halves, weight 0.3, v-prediction, KPL mse,
    --mixing object."""
    r, tr, bt = _compare(monkeypatch, 4, image_prior_weight=0.3, prediction_type="v_prediction", kpl_type="mse",
                         mixing="object")
    _check(r, tr)
    with pytest.raises(ValueError):
        tr.forward_backward(bt["latents"][:3], bt["noise"][:3], bt["timesteps"][:3], bt["input_ids"][:3],
                            bt["prior_ids"][:3])


@pytest.mark.long_cpu
def test_sd2x_shaped_step_on_cpu_matches_oracle(monkeypatch):
    """The SD-2.x branches at toy widths: linear (not 1x1-conv) Transformer2D projections, head_dim 64, and the OpenCLIP
    text encoder's erf-gelu, with v-prediction — what tests/test_gpu_step.py::test_step_sd21_openclip_h_vs_oracle covers
    at full width on the GPU."""
    from textboost_b200 import synthetic
    from textboost_b200.clip import ClipConfig
    from textboost_b200.unet import UNetConfig
    small = (UNetConfig(block_out_channels=(64, 128, 128, 128), attention_head_dim=(1, 2, 2, 2), cross_attention_dim=128,
                        use_linear_projection=True, sample_size=16),
             ClipConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2,
                        hidden_act="gelu"))
    monkeypatch.setattr(synthetic, "model_configs", lambda model: small)
    r, tr, _ = _compare(monkeypatch, 2, prediction_type="v_prediction", kpl_type="mse")
    _check(r, tr)


@pytest.mark.long_cpu
def test_three_steps_track_the_oracle(monkeypatch):
    """Optimiser state and the re-packed LoRA weights across steps: three product steps against three oracle steps from
    the same start.  Adam moves every parameter by ~lr per step whatever the gradient's size, so parameters may differ
    by a couple of lr where a gradient is ~0; the losses must keep tracking."""
    from oracle import harness, step_ref
    from textboost_b200 import synthetic
    engine_standin.install(monkeypatch)
    tr = synthetic.build_trainer("tiny", "cpu", seed=2, n_added=1, lora_b_std=0.02, keep_sd=True, learning_rate=1e-3,
                                 emb_learning_rate=1e-2)
    unet, te, te0, opt = harness.twin_of(tr, "cpu", torch.float32)
    V = tr.synthetic["clip_cfg"].vocab_size
    bt = synthetic.batch(2, 8, 5, V, "cpu")
    losses, refs = [], []
    for _ in range(3):
        losses.append(tr.step(bt["latents"], bt["noise"], bt["timesteps"], bt["input_ids"], bt["prior_ids"]).item())
        refs.append(step_ref.reference_step(unet, te, te0, bt["latents"], bt["noise"], bt["timesteps"], bt["input_ids"],
                                            bt["prior_ids"], n_base=V, kpl_weight=tr.kpl_weight, kpl_type="cos",
                                            optimizer=opt, max_grad_norm=tr.max_grad_norm,
                                            mean_norm=tr.mean_norm)["loss"].item())
    for a, b in zip(losses, refs):
        assert abs(a - b) < 5e-3 * abs(b), (losses, refs)
    assert tr.opt_state[4].item() == 3 and tr.opt_state[8].item() == 0
    st, r, worst = tr.te.state, tr.te.state.r, 0.0
    for l, lyr in enumerate(te.text_model.encoder.layers):
        for ti, t in enumerate(harness.LORA_TARGETS):
            m = getattr(lyr.self_attn, t)
            worst = max(worst, (st.A(l)[ti * r:(ti + 1) * r] - m.lora_A["default"].weight).abs().max().item(),
                        (st.B(l)[ti] - m.lora_B["default"].weight).abs().max().item())
    assert worst <= 3 * 2.1 * tr.lr, worst
    emb = te.get_input_embeddings().weight.detach()
    assert harness.rel_max(st.rows(), emb[V:]) < 5e-2


def test_lora_rank_zero_trains_only_the_added_rows(monkeypatch):
    """--lora_rank 0 (train_textboost.py:700-722): no adapter, frozen encoder, the added embedding rows are the only
    trainable state; nothing to clip (grad norm 0), frozen rows still decay (D8)."""
    r, tr, _ = _compare(monkeypatch, 2, kpl_type="cos", lora_r=0)
    assert tr.te.state.n_lora == 0 and tr.te.state.n_rows == 2
    assert abs(r["loss"] - r["loss_ref"]) < 2e-3 * abs(r["loss_ref"]) and r["pred_rel"] < 4e-3
    assert r["row_grad_rel"] < 5e-3
    assert r["grad_norm"] == 0.0 and r["grad_norm_ref"] == 0.0
    assert abs(r["added_norm"] - r["added_norm_ref"]) < 2e-3 * r["added_norm_ref"]
    # (the updated rows themselves are not compared elementwise: Adam's first step is lr * sign(g), so an element whose
    # gradient is ~0 may step the other way; the gradient and the post-step norm are what is pinned)
    assert abs(r["frozen_decay"] - r["frozen_decay_ref"]) < 1e-6


def test_gradient_accumulation_equals_one_step_over_the_joined_batch(monkeypatch):
    """--gradient_accumulation_steps 2 (accelerator.accumulate, train_textboost.py:1039; loss / 2 per micro-batch in
    accelerate's backward): two micro-batches of one image then ONE optimiser step  ==  one step over the two-image batch
    (both losses are batch means), with a linear warm-up schedule so the device-side --lr_scheduler is on the path."""
    from textboost_b200 import synthetic
    engine_standin.install(monkeypatch)
    kw = dict(seed=3, n_added=1, lora_b_std=0.02, learning_rate=1e-3, emb_learning_rate=1e-2, lr_scheduler="linear",
              lr_warmup_steps=2, max_train_steps=10)
    acc = synthetic.build_trainer("tiny", "cpu", gradient_accumulation_steps=2, **kw)
    one = synthetic.build_trainer("tiny", "cpu", **kw)
    V = acc.synthetic["clip_cfg"].vocab_size
    bt = synthetic.batch(2, 8, 7, V, "cpu")
    keys = ("latents", "noise", "timesteps", "input_ids", "prior_ids")
    for _ in range(2):  # two optimiser steps: the first runs at the warm-up's lr = 0, the second at lr / 2
        acc.step(*(bt[k][:1] for k in keys), sync_gradients=False)
        assert acc.opt_state[4].item() == one.opt_state[4].item()  # no optimiser step yet
        acc.step(*(bt[k][1:] for k in keys))
        one.step(*(bt[k] for k in keys))
    assert acc.opt_state[4].item() == 2 and abs(acc.opt_state[9].item() - 0.5) < 1e-6
    assert acc.opt.get_last_lr() == [pytest.approx(5e-3), pytest.approx(5e-4)]
    a, b = acc.te.state.params, one.te.state.params
    # Adam moves an element whose gradient is ~0 by +-lr whatever its size (fp16 noise picks the sign): at most two
    # steps of the second lr apart there, and equal to a hundredth of a step on average
    d = (a - b).abs()
    assert d.max().item() <= 2.1 * 5e-3 and d[:acc.te.state.n_lora].max().item() <= 2.1 * 5e-4
    assert d.mean().item() < 5e-6 and (d > 5e-5).float().mean().item() < 0.02
    rel = ((acc.opt.exp_avg - one.opt.exp_avg).norm() / one.opt.exp_avg.norm()).item()
    assert rel < 5e-3, rel  # fp16 activations: the one-image and two-image passes round differently


@pytest.mark.parametrize("name", ["small_quickgelu_qkv_r4", "small_gelu_qkvo_r8"])
def test_clip_engine_lora_on_cpu_matches_reference_golden(monkeypatch, name):
    """The product ClipEngine (LoRA fused into the QKV and out-projection GEMMs as K extensions; SIMT LoRA kernels
    from the product source) against the REFERENCE class run on merged weights (tests/golden/make_golden.py): the
    reference's q/k/v rank-4 configuration and the north star's "QKV/out projections" at rank 8, alpha 16."""
    import os
    import make_golden
    from textboost_b200.clip import ClipConfig, ClipEngine
    engine_standin.install(monkeypatch)
    case, targets, r, alpha, rows, glayers = make_golden.LORA_CASES[name]
    hidden, heads, layers, inter, act, n_added = make_golden.case_cfg(case)
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", f"clip_textboost_lora_{name}.pt"))
    sd = make_golden.make_weights(hidden, heads, layers, inter, n_added)
    sd.update(make_golden.lora_sd(make_golden.make_lora(hidden, layers, targets, r)))
    cfg = ClipConfig(hidden_size=hidden, intermediate_size=inter, num_hidden_layers=layers,
                     num_attention_heads=heads, hidden_act=act)
    eng = ClipEngine(cfg, sd, "cpu", lora_r=r, lora_alpha=alpha, n_base=make_golden.VOCAB, lora_targets=targets)
    assert eng.targets == tuple(targets) and eng.scaling == alpha / r
    ids, null, dout = make_golden.make_inputs(hidden, n_added)
    ids, dout = ids[:rows], dout[:rows]
    eng.set_null_embedding(null)
    eng.pack_lora()
    y = eng.forward(ids, save_for_backward=True)
    eng.state.grads.zero_()
    eng.backward(dout.clone())
    st = eng.state
    flat = torch.cat([t for l in range(layers) for ti in range(len(targets))
                      for t in (st.A(l, st.grads)[ti * r:(ti + 1) * r].flatten(), st.B(l, st.grads)[ti].flatten())])
    ref = gold["lora_grads_flat"]

    def rel(a, b):
        return ((a - b).abs().max() / b.abs().max()).item()
    assert rel(y, gold["out_fixed"]) < 2e-3
    assert rel(st.rows(st.grads), gold["grad_added_rows_fixed"]) < 3e-3
    assert ((flat - ref).norm() / ref.norm()).item() < 3e-3
    assert torch.nn.functional.cosine_similarity(flat, ref, dim=0).item() > 0.99999


@pytest.mark.long_cpu
def test_unet_cross_kv_lora_step_on_cpu_matches_oracle(monkeypatch):
    """--unet_params_to_train crossattn_kv (train_textboost.py:712-721, 838-841) through the product engines on the CPU,
    under the bf16 policy the mode is tied to: the bf16 host build of the SIMT sources (incl. csrc/unet_lora.cu), the
    trainer's third parameter group (optim.FlatAdamW: LoRA learning rate, not clipped), against the oracle step whose
    UNet carries the same adapters.  bf16 bounds = 8x the fp16 ones of _check."""
    from oracle import harness
    from textboost_b200 import synthetic
    engine_standin.install(monkeypatch, bf16=True)
    tr = synthetic.build_trainer("tiny", "cpu", seed=1, n_added=2, lora_b_std=0.02, keep_sd=True, learning_rate=1e-4,
                                 unet_lora_r=4)
    assert tr.opt_unet is not None and tr.opt_state[0].item() == 1.0  # no GradScaler under bf16
    V = tr.synthetic["clip_cfg"].vocab_size
    bt = synthetic.batch(2, 8, 3, V, "cpu")
    bt["input_ids"][1, 4] = V + 1
    before = tr.unet.kv_lora.params.clone()
    r = harness.compare_step(tr, bt)
    TX = 8.0
    assert abs(r["loss"] - r["loss_ref"]) < 2e-3 * TX * abs(r["loss_ref"])
    assert r["pred_rel"] < 4e-3 * TX
    assert r["lora_grad_rel_l2"] < 5e-3 * TX and r["row_grad_rel"] < 5e-3 * TX
    assert r["unet_lora_grad_norm_ref"] > 0
    assert r["unet_lora_grad_rel_l2"] < 5e-3 * TX and r["unet_lora_grad_cos"] > 1 - 1e-5 * TX ** 2
    assert r["unet_lora_param_max_abs_diff"] <= 2.1 * tr.lr
    assert (tr.unet.kv_lora.params - before).abs().max().item() > 0.5 * tr.lr
    assert torch.count_nonzero(tr.unet.kv_lora.grads) == 0 and tr.opt_unet.state[4].item() == 1
