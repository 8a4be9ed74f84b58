"""GPU probe: ClipEngine forward/backward vs the oracle TextBoostModelRef (fp32, same GPU)."""
import sys
import torch
sys.path.insert(0, ".")
from oracle import clip_ref  # noqa: E402
from textboost_b200 import clip as K  # noqa: E402

torch.manual_seed(0)
dev = "cuda"
torch.backends.cuda.matmul.allow_tf32 = False


def rel(a, b):
    return ((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-12)).item()


def run(name, rcfg, cfg, B, n_added=3):
    ref = clip_ref.TextBoostModelRef(rcfg)
    clip_ref.init_clip_(ref, seed=1)
    ref.resize_token_embeddings(rcfg.vocab_size + n_added)
    with torch.no_grad():
        ref.get_input_embeddings().weight[rcfg.vocab_size:] = ref.get_input_embeddings().weight[1000:1000 + n_added]
    ref.add_adapter(r=4)
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for lyr in ref.text_model.encoder.layers:
            for t in ("q_proj", "k_proj", "v_proj"):
                m = getattr(lyr.self_attn, t)
                m.lora_B["default"].weight.copy_(0.02 * torch.randn(m.lora_B["default"].weight.shape, generator=g))
    ref.get_input_embeddings().weight.requires_grad_(True)
    null = torch.randn(77, rcfg.hidden_size, generator=g)
    ref.set_null_embedding(null)
    ref = ref.to(dev)
    sd = {k: v for k, v in ref.state_dict().items()}
    eng = K.ClipEngine(cfg, sd, dev, lora_r=4, n_base=rcfg.vocab_size)
    eng.set_null_embedding(null)
    eng.pack_lora()
    ids = torch.full((B, 77), 49407, dtype=torch.int64)
    ids[:, 0] = 49406
    for b in range(B):
        n = 3 + b
        ids[b, 1:1 + n] = torch.randint(1000, 40000, (n,), generator=g)
        ids[b, 2] = rcfg.vocab_size + (b % n_added)
    ids[B - 1, 1:] = 49407  # empty prompt
    ids = ids.to(dev)
    dout = torch.randn(B, 77, rcfg.hidden_size, generator=g).to(dev)
    out = eng.forward(ids, save_for_backward=True)
    oref = ref(ids)
    print(f"[{name}] out rel={rel(out, oref):.3e}")
    eng.state.grads.zero_()
    eng.backward(dout.clone())
    oref.backward(dout)
    st = eng.state
    worst_a = worst_b = 0.0
    for l, lyr in enumerate(ref.text_model.encoder.layers):
        for ti, t in enumerate(("q_proj", "k_proj", "v_proj")):
            m = getattr(lyr.self_attn, t)
            ga = m.lora_A["default"].weight.grad
            gb = m.lora_B["default"].weight.grad
            worst_a = max(worst_a, rel(st.A(l, st.grads)[ti * 4:(ti + 1) * 4], ga))
            worst_b = max(worst_b, rel(st.B(l, st.grads)[ti], gb))
    ge = ref.get_input_embeddings().weight.grad[rcfg.vocab_size:]
    print(f"[{name}] dA rel={worst_a:.3e} dB rel={worst_b:.3e} d_rows rel={rel(st.rows(st.grads), ge):.3e} "
          f"(rows grad max {ge.abs().max().item():.3e})")


run("clip-l", clip_ref.ClipTextConfig.clip_l(), K.ClipConfig.clip_l(), 4)
run("openclip-h", clip_ref.ClipTextConfig.openclip_h(), K.ClipConfig.openclip_h(), 2)
