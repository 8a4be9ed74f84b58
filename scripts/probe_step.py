"""GPU probe: full TextBoost step (B200 kernels) vs oracle/step_ref (fp32 torch autograd on the same GPU)."""
import sys
import torch
sys.path.insert(0, ".")
from oracle import clip_ref, step_ref, unet_ref  # noqa: E402
from textboost_b200 import synthetic as S  # noqa: E402
from textboost_b200.clip import ClipConfig  # noqa: E402
from textboost_b200.unet import UNetConfig  # noqa: E402

dev = "cuda"
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def rel(a, b):
    return ((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-20)).item()


def check(B=2, n_added=2, kpl_type="cos", mixing=None):
    tr = S.build_trainer("sd15", dev, seed=42, n_added=n_added, lora_b_std=0.02, kpl_type=kpl_type,
                         mixing=mixing, learning_rate=1e-4)
    te, te0, unet = tr.te, tr.te0, tr.unet
    V = 49408
    bt = S.batch(B, 64, 7, V, dev)
    bt["input_ids"][1, 4] = V + 1
    # ---- oracle with identical weights
    ucfg = unet_ref.UNetConfig.sd15()
    runet = unet_ref.UNet2DConditionModelRef(ucfg).to(dev)
    runet.load_state_dict({k: v.float() for k, v in S.random_unet_sd(UNetConfig.sd15(), dev, 42).items()})
    runet.requires_grad_(False)
    csd = S.random_clip_sd(ClipConfig.clip_l(), V, dev, 43)
    rte0 = clip_ref.TextBoostModelRef(clip_ref.ClipTextConfig.clip_l()).to(dev)
    rte0.load_state_dict(csd, strict=False)
    rte0.set_null_embedding(te.null_embedding.clone())
    rte0.requires_grad_(False)
    rte = clip_ref.TextBoostModelRef(clip_ref.ClipTextConfig.clip_l()).to(dev)
    rte.load_state_dict(csd, strict=False)
    rte.resize_token_embeddings(V + n_added)
    with torch.no_grad():
        rte.get_input_embeddings().weight[V:] = te.state.rows()
    rte.set_null_embedding(te.null_embedding.clone())
    rte.add_adapter(r=4)
    rte = rte.to(dev)
    with torch.no_grad():
        for l, lyr in enumerate(rte.text_model.encoder.layers):
            for ti, t in enumerate(("q_proj", "k_proj", "v_proj")):
                m = getattr(lyr.self_attn, t)
                m.lora_A["default"].weight.copy_(te.state.A(l)[ti * 4:(ti + 1) * 4])
                m.lora_B["default"].weight.copy_(te.state.B(l)[ti])
    rte.get_input_embeddings().weight.requires_grad_(True)
    opt = step_ref.make_optimizer(rte, learning_rate=1e-4)
    ref = step_ref.reference_step(runet, rte, rte0, bt["latents"], bt["noise"], bt["timesteps"],
                                  bt["input_ids"], bt["prior_ids"], n_base=V, kpl_type=kpl_type,
                                  optimizer=opt, mixing=mixing, mean_norm=tr.mean_norm)
    # ---- ours
    scale = tr.opt_state[0].item()
    loss = tr.forward_backward(bt["latents"], bt["noise"], bt["timesteps"], bt["input_ids"], bt["prior_ids"])
    st = te.state
    g = st.grads.clone() / scale
    print(f"loss ours={loss.item():.6f} ref={ref['loss'].item():.6f}  pred rel={rel(tr._pred, ref['pred']):.3e}")
    wa = wb = 0.0
    ga_all, gr_all = [], []
    for l in range(te.nl):
        for ti, t in enumerate(("q_proj", "k_proj", "v_proj")):
            n = f"text_model.encoder.layers.{l}.self_attn.{t}."
            ra = ref["grad_lora"][n + "lora_A.default.weight"]
            rb = ref["grad_lora"][n + "lora_B.default.weight"]
            wa = max(wa, rel(st.A(l, g)[ti * 4:(ti + 1) * 4], ra))
            wb = max(wb, rel(st.B(l, g)[ti], rb))
            ga_all += [st.A(l, g)[ti * 4:(ti + 1) * 4].flatten(), st.B(l, g)[ti].flatten()]
            gr_all += [ra.flatten(), rb.flatten()]
    ga, gr = torch.cat(ga_all), torch.cat(gr_all)
    print(f"LoRA grads: worst per-tensor max-rel dA={wa:.3e} dB={wb:.3e}; global rel-L2={((ga - gr).norm() / gr.norm()).item():.3e} "
          f"cos={torch.nn.functional.cosine_similarity(ga, gr, dim=0).item():.6f}")
    print(f"row grads rel={rel(st.rows(g), ref['grad_rows']):.3e}  |g_rows|max={ref['grad_rows'].abs().max().item():.3e}")
    tr.all_reduce()
    tr.optimizer_step()
    torch.cuda.synchronize()
    print(f"grad norm ours={tr.opt_state[7].item():.5f} ref={ref['grad_norm'].item():.5f}  "
          f"added_norm ours={tr.added_norm.item():.5f} ref={ref['added_embedding_norm'].item():.5f}")
    rows_ref = rte.get_input_embeddings().weight[V:]
    print(f"rows after step rel={rel(st.rows(), rows_ref):.3e}; frozen-row decay ours={tr.opt_state[5].item():.8f} "
          f"ref={(rte.get_input_embeddings().weight[5] / csd['text_model.embeddings.token_embedding.weight'][5]).mean().item():.8f}")
    wa = 0.0
    for l, lyr in enumerate(rte.text_model.encoder.layers):
        for ti, t in enumerate(("q_proj", "k_proj", "v_proj")):
            m = getattr(lyr.self_attn, t)
            wa = max(wa, rel(st.A(l)[ti * 4:(ti + 1) * 4], m.lora_A["default"].weight),
                     rel(st.B(l)[ti], m.lora_B["default"].weight))
    print(f"LoRA params after step: worst rel={wa:.3e}", flush=True)
    del tr, runet, rte, rte0
    torch.cuda.empty_cache()


def timing(B=8):
    tr = S.build_trainer("sd15", dev, seed=42, n_added=1)
    bt = S.batch(B, 64, 7, 49408, dev)
    args = (bt["latents"], bt["noise"], bt["timesteps"], bt["input_ids"], bt["prior_ids"])
    for _ in range(3):
        tr.step(*args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(5):
        tr.step(*args)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"eager step B={B}: {ms:.2f} ms -> {B / ms * 1e3:.1f} img/s; loss={tr.loss.item():.5f} scale={tr.opt_state[0].item()} skipped={tr.opt_state[8].item()}")
    replay = tr.capture(*args)
    for _ in range(3):
        replay(*args)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        replay(*args)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"graph step B={B}: {ms:.2f} ms -> {B / ms * 1e3:.1f} img/s; loss={tr.loss.item():.5f} step={tr.opt_state[4].item()} skipped={tr.opt_state[8].item()}")
    print(f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")


check()
check(kpl_type="mse", mixing="object")
timing()
