"""GPU tests of the reference-facing Python surface: the TextBoostModel / UNet2DConditionModel mirrors driven the
way /root/reference/train_textboost.py drives the originals (encode_prompt -> unet(...).sample -> mse + kpl ->
scaled loss.backward()), against the reference's golden vectors and against the fused trainer; and the
train_textboost.py CLI end to end on a synthetic checkpoint in the diffusers layout."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
dev = "cuda"
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module", autouse=True)
def _need_gpu(built_lib):
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from textboost_b200 import _cabi
    _cabi.call("tb_check_device")


def relerr(a, b):
    return ((a.float().cpu() - b.float().cpu()).abs().max() / (b.float().abs().max().cpu() + 1e-12)).item()


def rel_l2(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-20)).item()


@pytest.mark.parametrize("case", ["small_quickgelu", "small_gelu"])
def test_textboost_model_mirror_matches_reference_golden(case):
    """The mirror class, used like the reference's (resize -> set_null_embedding -> model(ids)[0] ->
    .backward()), reproduces the outputs / added-row gradients of /root/reference's TextBoostModel."""
    import make_golden
    from textboost_b200.clip import ClipConfig
    from textboost_b200.text_encoder import TextBoostModel
    hidden, heads, layers, inter, act, n_added = make_golden.CASES[case]
    gold = torch.load(os.path.join(GOLDEN, f"clip_textboost_{case}.pt"))
    sd = make_golden.make_weights(hidden, heads, layers, inter, n_added)
    V = make_golden.VOCAB
    base = dict(sd)
    emb = sd["text_model.embeddings.token_embedding.weight"]
    base["text_model.embeddings.token_embedding.weight"] = emb[:V].clone()
    cfg = ClipConfig(vocab_size=V, hidden_size=hidden, intermediate_size=inter, num_hidden_layers=layers,
                     num_attention_heads=heads, hidden_act=act)
    te = TextBoostModel(cfg, base)
    te.resize_token_embeddings(V + n_added)
    te.get_input_embeddings().weight.data[V:] = emb[V:]
    ids, null, dout = make_golden.make_inputs(hidden, n_added)
    te.set_null_embedding(null)
    te.requires_grad_(False)
    te.get_input_embeddings().requires_grad_(True)
    te.to(dev)
    assert te.device.type == "cuda" and te.vocab_size == V + n_added
    out = te(ids.to(dev), return_dict=False)
    y = out[0]
    assert y.shape == (4, 77, hidden) and y.requires_grad
    assert relerr(y, gold["out_fixed"]) < 2e-3
    assert torch.equal(y[2].detach().cpu(), null)
    (y * dout.to(dev)).sum().backward()
    g = te.engine.state.rows(te.engine.state.grads)
    assert relerr(g, gold["grad_added_rows_fixed"]) < 3e-3
    # ModelOutput form + no-grad path give the same numbers
    with torch.no_grad():
        o2 = te(ids.to(dev))
    assert torch.equal(o2.last_hidden_state, y.detach()) and o2.pooler_output.shape == (4, hidden)
    # the dense embedding view == base rows | added rows
    w = te.get_input_embeddings().weight
    assert w.shape == (V + n_added, hidden) and torch.equal(w[V:].cpu(), emb[V:])


def test_reference_style_loop_through_mirrors_matches_fused_trainer():
    """train_textboost.py:1054-1108 written with the mirrors + autograd gives the gradients of the fused step."""
    from textboost_b200 import ops, synthetic
    from textboost_b200.lora import LoraConfig
    from textboost_b200.text_encoder import TextBoostModel
    from textboost_b200.unet_model import UNet2DConditionModel
    from textboost_b200.utils import encode_prompt
    import copy
    tr = synthetic.build_trainer("tiny", dev, seed=5, n_added=2, lora_b_std=0.02, keep_sd=True, kpl_weight=0.1)
    ucfg, ccfg = tr.synthetic["unet_cfg"], tr.synthetic["clip_cfg"]
    V = ccfg.vocab_size
    csd = {k: v.cpu() for k, v in tr.synthetic["clip_sd"].items()}
    te = TextBoostModel(ccfg, csd)
    te.set_null_embedding(tr.synthetic["null"])
    te0 = copy.deepcopy(te).eval().requires_grad_(False)
    te.resize_token_embeddings(V + 2)
    te.requires_grad_(False)
    te.add_adapter(LoraConfig(r=4, lora_alpha=4, init_lora_weights="gaussian",
                              target_modules=["q_proj", "k_proj", "v_proj"]))
    te.get_input_embeddings().requires_grad_(True)
    te.to(dev)
    te0.to(dev, dtype=torch.float16)
    te.engine.state.params.copy_(tr.te.state.params)  # same LoRA A/B and added rows as the trainer
    unet = UNet2DConditionModel(ucfg, {k: v.cpu() for k, v in tr.synthetic["unet_sd"].items()})
    unet.eval().requires_grad_(False)
    unet.to(dev, dtype=torch.float16)
    assert unet.dtype == torch.float16

    bt = synthetic.batch(3, 16, 11, V, dev)
    bt["input_ids"][1, 4] = V + 1
    bt["prior_ids"][2, 1:] = synthetic.EOS
    # fused trainer
    tr.forward_backward(bt["latents"], bt["noise"], bt["timesteps"], bt["input_ids"], bt["prior_ids"])
    scale = tr.opt_state[0].item()
    g_ref = tr.te.state.grads.clone() / scale
    loss_ref = tr.loss.item()
    # reference-style loop
    noisy, target = ops.add_noise(bt["latents"], bt["noise"], bt["timesteps"], tr.acp, False)
    ehs = encode_prompt(te, bt["input_ids"], None, text_encoder_use_attention_mask=False)
    pred = unet(noisy, bt["timesteps"], ehs.to(torch.float16)).sample
    loss = F.mse_loss(pred.float(), target.float(), reduction="none").mean()
    h = te(bt["prior_ids"])[0].float()
    with torch.no_grad():
        h0 = te0(bt["prior_ids"])[0].float()
    loss = loss + 0.1 * (1 - F.cosine_similarity(h, h0, dim=-1).mean())
    te.engine.state.grads.zero_()
    (loss * scale).backward()  # accelerate's GradScaler multiplies the loss before backward
    g = te.engine.state.grads / scale
    assert abs(loss.item() - loss_ref) < 1e-3 * abs(loss_ref)
    assert rel_l2(g, g_ref) < 5e-3
    names = dict(te.named_parameters())
    b0 = names["text_model.encoder.layers.0.self_attn.q_proj.lora_B.default.weight"]
    assert b0.grad is not None and b0.grad.shape == b0.shape and b0.grad.abs().sum() > 0
    # encoder parameters == the LoRA tensors (second optimiser group, train_textboost.py:835)
    assert sum(p.numel() for p in te.text_model.encoder.parameters() if p.requires_grad) == te.engine.state.n_lora


def test_cli_end_to_end_synthetic_checkpoint(tmp_path):
    """train_textboost.py on a random-init checkpoint in the diffusers layout: trains, writes the reference's
    output files, resumes from checkpoint-N."""
    import train_textboost as T
    from safetensors.torch import load_file
    from textboost_b200 import synthetic
    ck = str(tmp_path / "model")
    synthetic.write_pretrained(ck, "tiny", seed=2)
    out = str(tmp_path / "out")
    base = ["--pretrained_model_name_or_path", ck, "--output_dir", out, "--synthetic_data", "--resolution", "128",
            "--train_batch_size", "2", "--checkpointing_steps", "4", "--learning_rate", "1e-3",
            "--mixed_precision", "fp16", "--augment_inversion", "--log_every", "2"]
    loss = T.main(T.parse_args(base + ["--max_train_steps", "8"]))
    assert loss == loss and loss < 10
    files = set(os.listdir(out))
    assert {"text_encoder", "dog.bin", "checkpoint-4", "checkpoint-8", "training.log", "hflip.bin"} <= files
    row = torch.load(os.path.join(out, "dog.bin"))
    assert list(row) == ["<dog>"] and row["<dog>"].shape == (128,)
    assert torch.load(os.path.join(out, "hflip.bin"))["<hflip>"].shape == (1, 128)
    ad = load_file(os.path.join(out, "text_encoder", "adapter_model.safetensors"))
    kb = "base_model.model.text_model.encoder.layers.0.self_attn.v_proj.lora_B.weight"
    assert ad[kb].abs().sum() > 0  # B left its zero init: the LoRA trained
    assert {"dog.bin", "state.pt", "text_encoder"} <= set(os.listdir(os.path.join(out, "checkpoint-4")))
    # resume continues from step 8 to 10 without error and rotates nothing; this time the prior prompts come from a
    # human-written-prompts JSONL through the reference's prior pipeline (PriorPrompts + ShardedStream)
    import json
    jl = tmp_path / "prompts.jsonl"
    with open(jl, "w") as f:
        for i in range(9):
            f.write(json.dumps({"input": f"a photo of a thing {i}", "output": f"thing {i} in the snow"}) + "\n")
    loss2 = T.main(T.parse_args(base + ["--max_train_steps", "10", "--resume_from_checkpoint", "latest",
                                        "--prior_prompts_file", str(jl), "--class_token", "dog"]))
    assert loss2 == loss2
    assert T.RUN_INFO["prior_prompts"] == 18
    # modes outside the built path fail loudly rather than silently training something else
    with pytest.raises(NotImplementedError):
        T.main(T.parse_args(base + ["--unet_params_to_train", "crossattn_kv"]))
    # --lora_rank 0: textual inversion only (train_textboost.py:700, 722, 1237): the embeddings train, no adapter is
    # written; with a cosine schedule over the run and two accumulated micro-batches per optimiser step
    out0 = str(tmp_path / "out0")
    base0 = [a if a != out else out0 for a in base]
    loss0 = T.main(T.parse_args(base0 + ["--lora_rank", "0", "--max_train_steps", "6", "--lr_scheduler", "cosine",
                                         "--lr_warmup_steps", "2", "--gradient_accumulation_steps", "2"]))
    assert loss0 == loss0 and "text_encoder" not in os.listdir(out0) and "dog.bin" in os.listdir(out0)
    assert torch.load(os.path.join(out0, "dog.bin"))["<dog>"].shape == (128,)


def test_cli_trains_from_images_through_vae_front_end(tmp_path):
    """The reference's real data path (train_textboost.py:856-890, 1027-1037): instance images -> PairedAugmentation
    -> TextBoostDataset -> pixel_values -> AutoencoderKL encoder on the GPU -> latents -> the fused step."""
    import json
    import make_augment_golden as G
    import train_textboost as T
    from textboost_b200 import synthetic
    ck = str(tmp_path / "model")
    synthetic.write_pretrained(ck, "tiny", seed=3, vae_channels=(64, 64, 128, 128))
    imgs = tmp_path / "dog"
    imgs.mkdir()
    for i, size in enumerate([(160, 140), (128, 128), (150, 200)]):
        G.make_image(size, i).save(imgs / f"{i}.png")
    jl = tmp_path / "prompts.jsonl"
    with open(jl, "w") as f:
        for i in range(6):
            f.write(json.dumps({"input": f"a photo of a thing {i}", "output": "NONE"}) + "\n")
    out = str(tmp_path / "out")
    loss = T.main(T.parse_args([
        "--pretrained_model_name_or_path", ck, "--output_dir", out, "--instance_data_dir", str(imgs),
        "--resolution", "128", "--train_batch_size", "2", "--max_train_steps", "6", "--learning_rate", "1e-3",
        "--mixed_precision", "fp16", "--augment", "pda", "--augment_inversion", "--template", "textboost",
        "--prior_prompts_file", str(jl), "--class_token", "dog", "--log_every", "2", "--seed", "11"]))
    assert loss == loss and loss < 10
    assert T.RUN_INFO["instance_images"] == 3 and T.RUN_INFO["prior_prompts"] == 6
    assert {"text_encoder", "dog.bin", "hflip.bin", "training.log"} <= set(os.listdir(out))
    with pytest.raises(ValueError):  # no data source at all
        T.main(T.parse_args(["--pretrained_model_name_or_path", ck, "--output_dir", out, "--max_train_steps", "1"]))
