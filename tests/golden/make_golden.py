"""Generates tests/golden/clip_textboost_*.pt by running the REFERENCE's own class
(/root/reference/textboost/text_encoder.py::TextBoostModel, on the installed transformers 5.5.0)
in this container.  /root/reference does not exist on the GPU box, so only the outputs travel.

Weights are not stored: both this script and the tests rebuild them from a seed with `make_weights`
(torch CPU generator, bit-reproducible), and feed them to the reference / the oracle / the CUDA engine.

    python tests/golden/make_golden.py
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (hidden, heads, layers, intermediate, act, n_added)
    "small_quickgelu": (128, 2, 2, 256, "quick_gelu", 2),
    "small_gelu": (64, 1, 3, 192, "gelu", 1),
}
VOCAB = 49408
L = 77


def make_weights(hidden, heads, layers, inter, n_added, seed=1234):
    """HF-keyed fp32 state dict, deterministic.  Projection / MLP weights ~ N(0, 0.08) at the toy widths and
    N(0, 0.03) at full width (per-layer gain 0.08 sqrt(128) ~ 0.03 sqrt(768) ~ 0.9: a 12-layer random net with gain
    2.2 per layer is ill-conditioned and would test fp16 error amplification, not the encoder)."""
    wstd = 0.08 if hidden <= 256 else 0.03
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def rn(*shape, std=0.02):
        return torch.randn(*shape, generator=g) * std

    sd["text_model.embeddings.token_embedding.weight"] = rn(VOCAB + n_added, hidden)
    sd["text_model.embeddings.position_embedding.weight"] = rn(L, hidden)
    for l in range(layers):
        p = f"text_model.encoder.layers.{l}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            sd[p + f"self_attn.{n}.weight"] = rn(hidden, hidden, std=wstd)
            sd[p + f"self_attn.{n}.bias"] = rn(hidden)
        for n in ("layer_norm1", "layer_norm2"):
            sd[p + n + ".weight"] = 1.0 + rn(hidden, std=0.1)
            sd[p + n + ".bias"] = rn(hidden)
        sd[p + "mlp.fc1.weight"] = rn(inter, hidden, std=wstd)
        sd[p + "mlp.fc1.bias"] = rn(inter)
        sd[p + "mlp.fc2.weight"] = rn(hidden, inter, std=wstd if hidden <= 256 else wstd / 2)
        sd[p + "mlp.fc2.bias"] = rn(hidden)
    sd["text_model.final_layer_norm.weight"] = 1.0 + rn(hidden, std=0.1)
    sd["text_model.final_layer_norm.bias"] = rn(hidden)
    return sd


# LoRA goldens: the reference class has no peft here, but peft's LoRA Linear is linear algebra on the weight:
# y = x (W + s B A)^T + b, so the REFERENCE class run on merged weights W' gives the LoRA forward, and
#   dL/dA = s B^T (dL/dW'),   dL/dB = s (dL/dW') A^T
# gives the LoRA gradients from the reference's own autograd (train_textboost.py:702-709 configures r, alpha,
# targets; peft tuners/lora/layer.py Linear.forward is the arithmetic).
QKV = ("q_proj", "k_proj", "v_proj")
LORA_CASES = {
    # name: (weights case, targets, r, lora_alpha, batch rows of make_inputs used, layers whose LoRA gradients
    #        are stored -- None = all; the rank-16 full-size case keeps three layers so the file stays ~1.6 MB)
    "small_quickgelu_qkv_r4": ("small_quickgelu", QKV, 4, 4, 4, None),
    "small_gelu_qkvo_r8": ("small_gelu", QKV + ("out_proj",), 8, 16, 4, None),
    "clip_l_qkv_r4": ("clip_l", QKV, 4, 4, 2, None),    # BASELINE.json configs[0]: CLIP-L, rank 4, bs 2, 77 tokens
    "clip_l_qkvo_r16": ("clip_l", QKV + ("out_proj",), 16, 16, 2, (0, 6, 11)),
}
FULL_CASES = {"clip_l": (768, 12, 12, 3072, "quick_gelu", 2)}


def case_cfg(case):
    return CASES[case] if case in CASES else FULL_CASES[case]


def make_lora(hidden, layers, targets, r, seed=77):
    """{(layer, target): (A [r, hidden], B [hidden, r])}: A ~ N(0, 1/r) as peft's gaussian init, B ~ N(0, 0.02)
    (non-zero so that dL/dA is not identically zero)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for l in range(layers):
        for t in targets:
            out[(l, t)] = (torch.randn(r, hidden, generator=g) / r, 0.02 * torch.randn(hidden, r, generator=g))
    return out


def lora_sd(lora):
    """the same factors under peft's state-dict names (what ClipEngine / the oracle load)."""
    sd = {}
    for (l, t), (A, B) in lora.items():
        p = f"text_model.encoder.layers.{l}.self_attn.{t}."
        sd[p + "lora_A.default.weight"], sd[p + "lora_B.default.weight"] = A, B
    return sd


def make_inputs(hidden, n_added, seed=99):
    g = torch.Generator().manual_seed(seed)
    ids = torch.full((4, L), 49407, dtype=torch.int64)
    ids[:, 0] = 49406
    ids[0, 1:5] = torch.tensor([320, VOCAB, 1929, 525])            # 'a <new0> dog ...'
    ids[1, 1:9] = torch.randint(1000, 40000, (8,), generator=g)
    ids[1, 3] = VOCAB + n_added - 1
    # row 2: empty prompt (EOS right after BOS) -> whole row overwritten by the null embedding
    ids[3, 1:20] = torch.randint(1000, 40000, (19,), generator=g)
    null = torch.randn(L, hidden, generator=g)
    dout = torch.randn(4, L, hidden, generator=g)
    return ids, null, dout


def main():
    sys.path.insert(0, "/root/reference")
    from transformers import CLIPTextConfig
    from textboost.text_encoder import TextBoostModel  # the reference class

    for name, (hidden, heads, layers, inter, act, n_added) in CASES.items():
        cfg = CLIPTextConfig(vocab_size=VOCAB + n_added, hidden_size=hidden, intermediate_size=inter,
                             num_hidden_layers=layers, num_attention_heads=heads,
                             max_position_embeddings=L, hidden_act=act, projection_dim=hidden,
                             bos_token_id=49406, eos_token_id=49407, pad_token_id=1)
        cfg._attn_implementation = "eager"
        model = TextBoostModel(cfg).eval()
        sd = make_weights(hidden, heads, layers, inter, n_added)
        missing, unexpected = model.load_state_dict(sd, strict=False)
        assert not unexpected, unexpected
        assert all(("null_embedding" in m) or ("position_ids" in m) for m in missing), missing
        ids, null, dout = make_inputs(hidden, n_added)
        out = {}
        for fixed in (False, True):
            m = TextBoostModel(cfg).eval()
            m.load_state_dict(sd, strict=False)
            if fixed:
                m.set_null_embedding(null.clone())
            else:
                m.null_embedding = null.clone()   # buffer set without enabling the position-0 override
            emb = m.get_input_embeddings().weight
            emb.requires_grad_(True)
            y = m(ids, return_dict=False)[0]
            (y * dout).sum().backward()
            key = "fixed" if fixed else "plain"
            out[f"out_{key}"] = y.detach().clone()
            out[f"grad_added_rows_{key}"] = emb.grad[VOCAB:].detach().clone()
            out[f"grad_row_320_{key}"] = emb.grad[320].detach().clone()
        out["meta"] = {"case": name, "cfg": (hidden, heads, layers, inter, act, n_added),
                       "transformers": __import__("transformers").__version__,
                       "torch": str(torch.__version__), "reference": "textboost/text_encoder.py:17-87"}
        path = os.path.join(HERE, f"clip_textboost_{name}.pt")
        torch.save(out, path)
        print("wrote", path, os.path.getsize(path), "bytes")

    for name, (case, targets, r, alpha, rows, glayers) in LORA_CASES.items():
        hidden, heads, layers, inter, act, n_added = case_cfg(case)
        cfg = CLIPTextConfig(vocab_size=VOCAB + n_added, hidden_size=hidden, intermediate_size=inter,
                             num_hidden_layers=layers, num_attention_heads=heads,
                             max_position_embeddings=L, hidden_act=act, projection_dim=hidden,
                             bos_token_id=49406, eos_token_id=49407, pad_token_id=1)
        cfg._attn_implementation = "eager"
        sd = make_weights(hidden, heads, layers, inter, n_added)
        lora = make_lora(hidden, layers, targets, r)
        s = alpha / r
        merged = dict(sd)
        for (l, t), (A, B) in lora.items():
            k = f"text_model.encoder.layers.{l}.self_attn.{t}.weight"
            merged[k] = sd[k] + s * (B @ A)
        m = TextBoostModel(cfg).eval()
        m.load_state_dict(merged, strict=False)
        ids, null, dout = make_inputs(hidden, n_added)
        ids, dout = ids[:rows], dout[:rows]
        m.set_null_embedding(null.clone())
        m.requires_grad_(False)
        emb = m.get_input_embeddings().weight
        emb.requires_grad_(True)
        ws = {}
        for (l, t) in lora:
            w = getattr(m.text_model.encoder.layers[l].self_attn, t).weight
            w.requires_grad_(True)
            ws[(l, t)] = w
        y = m(ids, return_dict=False)[0]
        (y * dout).sum().backward()
        flat = []
        for (l, t), (A, B) in lora.items():      # (layer, target) order, A then B
            if glayers is not None and l not in glayers:
                continue
            dW = ws[(l, t)].grad
            flat += [(s * B.t() @ dW).flatten(), (s * dW @ A.t()).flatten()]
        out = {"out_fixed": y.detach().clone(), "grad_added_rows_fixed": emb.grad[VOCAB:].detach().clone(),
               "lora_grads_flat": torch.cat(flat),
               "meta": {"case": case, "targets": targets, "r": r, "lora_alpha": alpha, "rows": rows, "grad_layers": glayers,
                        "transformers": __import__("transformers").__version__, "torch": str(torch.__version__),
                        "reference": "textboost/text_encoder.py:17-87 on merged weights W + (alpha/r) B A",
                        "order": "for layer: for target: dA [r, D] then dB [D, r]"}}
        path = os.path.join(HERE, f"clip_textboost_lora_{name}.pt")
        torch.save(out, path)
        print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
