"""CPU tests of the oracle (test infrastructure) against the reference-generated golden vectors and the
closed forms / identities that pin each restated piece."""
import os

import pytest
import torch

from oracle import clip_ref, ddpm_ref, step_ref, unet_ref

import make_golden  # tests/golden/make_golden.py (weights/inputs are rebuilt from seeds)

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def build_oracle_clip(case):
    hidden, heads, layers, inter, act, n_added = make_golden.case_cfg(case)
    cfg = clip_ref.ClipTextConfig(vocab_size=make_golden.VOCAB + n_added, hidden_size=hidden,
                                  intermediate_size=inter, num_hidden_layers=layers,
                                  num_attention_heads=heads, hidden_act=act)
    m = clip_ref.TextBoostModelRef(cfg)
    sd = make_golden.make_weights(hidden, heads, layers, inter, n_added)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and missing == ["null_embedding"]
    return m, sd, cfg


@pytest.mark.parametrize("case", list(make_golden.CASES))
@pytest.mark.parametrize("fixed", [False, True])
def test_clip_oracle_matches_reference_golden(case, fixed):
    """oracle/clip_ref.py == /root/reference TextBoostModel (transformers 5.5.0, eager) on the same weights."""
    gold = torch.load(os.path.join(GOLDEN, f"clip_textboost_{case}.pt"))
    m, sd, cfg = build_oracle_clip(case)
    n_added = make_golden.CASES[case][5]
    ids, null, dout = make_golden.make_inputs(cfg.hidden_size, n_added)
    if fixed:
        m.set_null_embedding(null.clone())
    else:
        m.null_embedding = null.clone()
    emb = m.get_input_embeddings().weight
    emb.requires_grad_(True)
    y = m(ids)
    (y * dout).sum().backward()
    key = "fixed" if fixed else "plain"
    torch.testing.assert_close(y, gold[f"out_{key}"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(emb.grad[make_golden.VOCAB:], gold[f"grad_added_rows_{key}"], rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(emb.grad[320], gold[f"grad_row_320_{key}"], rtol=1e-4, atol=1e-6)
    # text_encoder.py:71-86: the empty prompt row is the null embedding, exactly
    assert torch.equal(y[2], null)
    if fixed:
        assert torch.equal(y[:, 0], null[0].expand(4, -1))


@pytest.mark.parametrize("name", list(make_golden.LORA_CASES))
def test_clip_lora_oracle_matches_reference_golden(name):
    """oracle/clip_ref.py's peft-LoRA restatement (LoraLinear) against the REFERENCE class run on merged weights
    W + (alpha/r) B A (tests/golden/make_golden.py): output, dA / dB (derived from the reference's dL/dW'), and the
    added-row embedding gradients.  Pins SURVEY.md §8 a4 to reference-run outputs, incl. full-size CLIP-L at
    BASELINE.json configs[0] (rank 4, bs 2, 77 tokens) and the out_proj / rank-16 variants."""
    case, targets, r, alpha, rows, glayers = make_golden.LORA_CASES[name]
    gold = torch.load(os.path.join(GOLDEN, f"clip_textboost_lora_{name}.pt"))
    m, sd, cfg = build_oracle_clip(case)
    n_added = make_golden.case_cfg(case)[5]
    lora = make_golden.make_lora(cfg.hidden_size, cfg.num_hidden_layers, targets, r)
    m.requires_grad_(False)
    m.add_adapter(r=r, lora_alpha=alpha, target_modules=targets)
    missing, unexpected = m.load_state_dict(
        make_golden.lora_sd(lora), strict=False)
    assert not unexpected
    ids, null, dout = make_golden.make_inputs(cfg.hidden_size, n_added)
    ids, dout = ids[:rows], dout[:rows]
    m.set_null_embedding(null.clone())
    emb = m.get_input_embeddings().weight
    emb.requires_grad_(True)
    y = m(ids)
    (y * dout).sum().backward()
    # merged (W + sBA rounded to fp32) vs factored arithmetic: fp32 rounding through up to 12 layers
    torch.testing.assert_close(y, gold["out_fixed"], rtol=1e-4, atol=1e-5 * gold["out_fixed"].abs().max().item())
    gr = gold["grad_added_rows_fixed"]
    torch.testing.assert_close(emb.grad[make_golden.VOCAB:], gr, rtol=1e-4, atol=5e-5 * gr.abs().max().item())
    flat = []
    for (l, t) in lora:
        if glayers is not None and l not in glayers:
            continue
        mod = getattr(m.text_model.encoder.layers[l].self_attn, t)
        flat += [mod.lora_A["default"].weight.grad.flatten(), mod.lora_B["default"].weight.grad.flatten()]
    flat = torch.cat(flat)
    ref = gold["lora_grads_flat"]
    assert flat.shape == ref.shape
    assert ((flat - ref).norm() / ref.norm()).item() < 5e-5  # fp32 rounding (merged vs factored weights)
    torch.testing.assert_close(flat, ref, rtol=1e-3, atol=5e-5 * ref.abs().max().item())


def test_lora_restatement_equals_merged_weights():
    """peft LoRA Linear: base(x) + (alpha/r) B A x == x (W + (alpha/r) B A)^T + b."""
    torch.manual_seed(0)
    base = torch.nn.Linear(48, 32)
    l = clip_ref.LoraLinear(base, r=4, lora_alpha=8)
    assert l.scaling == 2.0
    assert torch.count_nonzero(l.lora_B["default"].weight) == 0  # init_lora_weights="gaussian": B = 0
    assert abs(l.lora_A["default"].weight.std().item() - 0.25) < 0.05  # A ~ N(0, (1/r)^2)
    with torch.no_grad():
        l.lora_B["default"].weight.normal_()
    x = torch.randn(5, 48)
    merged = base.weight + 2.0 * l.lora_B["default"].weight @ l.lora_A["default"].weight
    torch.testing.assert_close(l(x), x @ merged.t() + base.bias, rtol=1e-5, atol=1e-5)
    assert not base.weight.requires_grad and l.lora_A["default"].weight.requires_grad


def test_add_adapter_trains_only_lora():
    m, _, _ = build_oracle_clip("small_gelu")
    m.add_adapter(r=4)
    names = [n for n, p in m.named_parameters() if p.requires_grad]
    assert names and all("lora_" in n for n in names)
    assert len(names) == 3 * 2 * 3  # layers x (A,B) x (q,k,v)


def test_ddpm_constants():
    """SURVEY.md §4 probed values."""
    acp = ddpm_ref.alphas_cumprod()
    assert abs(acp[0].item() - 0.99915) < 1e-5
    assert abs(acp[499].item() - 0.27767) < 1e-4
    assert abs(acp[999].item() - 0.00466) < 1e-5
    p = ddpm_ref.timestep_probs()
    assert p[0].item() == 0.0 and abs(p[999].item() - 1.547e-3) < 2e-6
    assert abs(p.sum().item() - 1) < 1e-5
    assert abs((p * torch.arange(1000)).sum().item() - 584.3) < 0.2
    assert torch.all(p[1:] >= p[:-1])
    x0, eps = torch.randn(3, 4, 8, 8), torch.randn(3, 4, 8, 8)
    t = torch.tensor([0, 499, 999])
    xt = ddpm_ref.add_noise(x0, eps, t)
    v = ddpm_ref.get_velocity(x0, eps, t)
    sa, sb = acp[t].sqrt().view(3, 1, 1, 1), (1 - acp[t]).sqrt().view(3, 1, 1, 1)
    # (x_t, v) is a rotation of (x0, eps): x0 = sa x_t - sb v ; eps = sb x_t + sa v
    torch.testing.assert_close(sa * xt - sb * v, x0, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(sb * xt + sa * v, eps, rtol=1e-4, atol=1e-5)


def test_unet_oracle_structure():
    from textboost_b200 import synthetic
    from textboost_b200.unet import UNetConfig
    with torch.device("meta"):
        m = unet_ref.UNet2DConditionModelRef(unet_ref.UNetConfig.sd15())
    sd = m.state_dict()
    assert sum(v.numel() for v in sd.values()) == 859_520_964  # SD-1.5 UNet parameter count
    assert len(sd) == 686
    shapes = synthetic.unet_shapes(UNetConfig.sd15())
    assert {k: tuple(v.shape) for k, v in sd.items()} == shapes
    with torch.device("meta"):
        m2 = unet_ref.UNet2DConditionModelRef(unet_ref.UNetConfig.sd21())
    assert sum(v.numel() for v in m2.state_dict().values()) == 865_910_724  # SD-2.1 UNet
    assert {k: tuple(v.shape) for k, v in m2.state_dict().items()} == synthetic.unet_shapes(UNetConfig.sd21())


def test_unet_oracle_tiny_forward_backward():
    cfg = unet_ref.UNetConfig.tiny()
    m = unet_ref.init_unet_(unet_ref.UNet2DConditionModelRef(cfg)).requires_grad_(False)
    x = torch.randn(2, 4, 16, 16)
    ehs = torch.randn(2, 77, cfg.cross_attention_dim, requires_grad=True)
    y = m(x, torch.tensor([3, 800]), ehs)
    assert y.shape == x.shape and torch.isfinite(y).all() and 0.1 < y.std() < 10
    y.square().mean().backward()
    assert ehs.grad.abs().max() > 0
    # samples are independent: the gradient wrt ehs[1] does not depend on sample 0
    ehs2 = ehs.detach().clone().requires_grad_(True)
    m(x[1:], torch.tensor([800]), ehs2[1:]).square().mean().backward()
    torch.testing.assert_close(ehs2.grad[1] * 0.5, ehs.grad[1], rtol=1e-3, atol=1e-7)


def test_timestep_embedding_layout():
    e = unet_ref.timestep_embedding(torch.tensor([0, 10]), 320)
    assert e.shape == (2, 320)
    assert torch.allclose(e[0, :160], torch.ones(160)) and torch.allclose(e[0, 160:], torch.zeros(160))  # cos first
    assert abs(e[1, 160].item() - torch.sin(torch.tensor(10.0)).item()) < 1e-6


def _tiny_step_models(seed=0, n_added=2, hidden=64):
    ucfg = unet_ref.UNetConfig.tiny(cross_attention_dim=hidden)
    unet = unet_ref.init_unet_(unet_ref.UNet2DConditionModelRef(ucfg), seed).requires_grad_(False)
    ccfg = clip_ref.ClipTextConfig(hidden_size=hidden, intermediate_size=128, num_hidden_layers=2,
                                   num_attention_heads=1)
    te0 = clip_ref.init_clip_(clip_ref.TextBoostModelRef(ccfg), seed + 1)
    null = torch.randn(77, hidden, generator=torch.Generator().manual_seed(seed + 2))
    te0.set_null_embedding(null)
    import copy
    te = copy.deepcopy(te0)
    te0.requires_grad_(False)
    te.resize_token_embeddings(ccfg.vocab_size + n_added)
    with torch.no_grad():
        te.get_input_embeddings().weight[ccfg.vocab_size:] = te.get_input_embeddings().weight[500:500 + n_added]
    te.add_adapter(r=4)
    with torch.no_grad():
        for n, p in te.named_parameters():
            if "lora_B" in n:
                p.normal_(std=0.02)
    te.get_input_embeddings().weight.requires_grad_(True)
    return unet, te, te0, ccfg


def _tiny_batch(B, V, seed=5):
    g = torch.Generator().manual_seed(seed)
    ids = torch.full((B, 77), 49407, dtype=torch.int64)
    ids[:, 0] = 49406
    ids[:, 1], ids[:, 2], ids[:, 3] = 320, V, 1929
    pr = torch.full((B, 77), 49407, dtype=torch.int64)
    pr[:, 0] = 49406
    pr[: B - 1, 1:6] = torch.randint(1000, 40000, (B - 1, 5), generator=g)  # last prior prompt is empty
    return (torch.randn(B, 4, 16, 16, generator=g), torch.randn(B, 4, 16, 16, generator=g),
            torch.randint(0, 1000, (B,), generator=g), ids, pr)


def test_reference_step_semantics():
    """Parity traps of SURVEY.md Appendix E on the oracle step itself."""
    unet, te, te0, ccfg = _tiny_step_models()
    V = ccfg.vocab_size
    lat, noise, t, ids, pr = _tiny_batch(3, V)
    emb = te.get_input_embeddings().weight
    w0 = emb.detach().clone()
    mean_norm = emb.norm(dim=-1).mean().item()
    opt = step_ref.make_optimizer(te, learning_rate=1e-4, emb_learning_rate=1e-3)
    out = step_ref.reference_step(unet, te, te0, lat, noise, t, ids, pr, n_base=V, optimizer=opt,
                                  mean_norm=mean_norm)
    assert out["grad_rows"].abs().max() > 0 and out["d_ehs"].abs().max() > 0
    assert all(g.abs().max() > 0 for g in out["grad_lora"].values())
    # D8: decoupled weight decay shrinks every frozen row by (1 - lr*wd) even with zero gradient
    torch.testing.assert_close(emb[:V].detach(), w0[:V] * (1 - 1e-3 * 1e-2), rtol=1e-6, atol=0)
    # the placeholder row (id V) moved, the unused added row (V+1) only decayed (its Adam update is 0)
    assert (emb[V].detach() - w0[V]).abs().max() > 1e-5
    # renorm only ever shrinks
    assert emb[V:].norm(dim=-1).max() <= max(mean_norm, w0[V:].norm(dim=-1).max()) + 1e-6
    # Appendix E.1: an all-empty-prompt batch gives exactly zero gradient through the instance path
    ids_empty = torch.full_like(ids, 49407)
    ids_empty[:, 0] = 49406
    out2 = step_ref.reference_step(unet, te, te0, lat, noise, t, ids_empty, None, n_base=V, kpl_weight=0.0)
    assert out2["d_ehs"].abs().max() > 0          # the UNet still sends a gradient back ...
    assert all(g.abs().max() == 0 for g in out2["grad_lora"].values())   # ... the override blocks it


def test_kpl_zero_at_overridden_slots():
    _, te, te0, ccfg = _tiny_step_models()
    pr = _tiny_batch(3, ccfg.vocab_size)[4]
    h, h0 = te(pr), te0(pr)
    cos = torch.nn.functional.cosine_similarity(h, h0, dim=-1)
    assert torch.all(cos[:, 0] == 1) or torch.allclose(cos[:, 0], torch.ones(3), atol=1e-6)  # position 0 fixed
    assert torch.allclose(cos[2], torch.ones(77), atol=1e-6)  # empty prompt row == null embedding in both


def test_lazy_decay_equals_dense_adamw():
    """AdamW over the whole embedding == scalar decay on frozen rows + AdamW on live rows."""
    torch.manual_seed(0)
    V, n, D, lr, wd = 50, 3, 8, 1e-3, 1e-2
    w = torch.nn.Parameter(torch.randn(V + n, D))
    w0 = w.detach().clone()
    opt = torch.optim.AdamW([w], lr=lr, weight_decay=wd)
    rows = w0[V:].clone()
    m = torch.zeros(n, D)
    v = torch.zeros(n, D)
    c = 1.0
    for step in range(1, 6):
        g = torch.zeros(V + n, D)
        g[V:] = torch.randn(n, D)
        w.grad = g.clone()
        opt.step()
        rows = rows * (1 - lr * wd)
        m = 0.9 * m + 0.1 * g[V:]
        v = 0.999 * v + 0.001 * g[V:] ** 2
        rows = rows - lr / (1 - 0.9 ** step) * m / (v.sqrt() / (1 - 0.999 ** step) ** 0.5 + 1e-8)
        c *= (1 - lr * wd)
    torch.testing.assert_close(w.detach()[:V], w0[:V] * c, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(w.detach()[V:], rows, rtol=1e-5, atol=1e-6)


def test_vae_encoder_oracle_structure():
    """oracle/vae_ref.py (image half of SURVEY.md §8 f1, groundwork): diffusers key names, the published encoder
    parameter count of the SD VAE (34,163,592 + 72 for quant_conv), latent geometry and the sampling formula."""
    from oracle import vae_ref
    m = vae_ref.AutoencoderKLEncoderRef()
    keys = set(m.state_dict())
    assert sum(p.numel() for p in m.encoder.parameters()) == 34_163_592
    assert sum(p.numel() for p in m.quant_conv.parameters()) == 72
    for k in ("encoder.conv_in.weight", "encoder.down_blocks.1.resnets.0.conv_shortcut.weight",
              "encoder.down_blocks.2.downsamplers.0.conv.bias", "encoder.mid_block.attentions.0.to_out.0.weight",
              "encoder.mid_block.attentions.0.group_norm.weight", "encoder.conv_norm_out.bias", "quant_conv.weight"):
        assert k in keys, k
    assert "encoder.down_blocks.3.downsamplers.0.conv.weight" not in keys
    small = vae_ref.AutoencoderKLEncoderRef(vae_ref.VAEConfig(block_out_channels=(32, 32, 64, 64)))
    x = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(0))
    eps = torch.randn(2, 4, 8, 8, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        mean, std = small.moments(x)
        z = small.encode_latents(x, eps)
    assert mean.shape == (2, 4, 8, 8) and (std > 0).all()
    torch.testing.assert_close(z, (mean + std * eps) * 0.18215)


def test_image_prior_loss_semantics():
    """--with_image_prior (train_textboost.py:1077-1094): the batch is [instance | class] halves and
    loss = mse(instance) + image_ppl_weight * mse(class); with weight 1 that is twice the plain mean over the batch,
    with weight 0 the class half gets no gradient at all."""
    unet, te, te0, ccfg = _tiny_step_models()
    V = ccfg.vocab_size
    lat, noise, t, ids, pr = _tiny_batch(4, V)
    plain, pred, _ = step_ref.forward_loss(unet, te, te0, lat, noise, t, ids, None, kpl_weight=0.0)
    both, _, _ = step_ref.forward_loss(unet, te, te0, lat, noise, t, ids, None, kpl_weight=0.0, image_ppl_weight=1.0)
    torch.testing.assert_close(both, 2 * plain, rtol=1e-5, atol=0)
    w = 0.3
    mixed, _, _ = step_ref.forward_loss(unet, te, te0, lat, noise, t, ids, None, kpl_weight=0.0, image_ppl_weight=w)
    want = ((pred[:2] - noise[:2]) ** 2).mean() + w * ((pred[2:] - noise[2:]) ** 2).mean()
    torch.testing.assert_close(mixed, want, rtol=1e-5, atol=0)
    out = step_ref.reference_step(unet, te, te0, lat, noise, t, ids, None, n_base=V, kpl_weight=0.0,
                                  image_ppl_weight=0.0)
    assert out["d_ehs"][:2].abs().max() > 0 and out["d_ehs"][2:].abs().max() == 0
