"""GPU library baseline: the reference step as the pinned libraries would execute it on this
B200 if their version pins were lifted — the oracle modules (diffusers-0.29 UNet graph, transformers CLIP, peft
LoRA semantics) in fp16 through torch 2.11: cuDNN convolutions, cuBLASLt linears, torch SDPA attention (what
diffusers' AttnProcessor2_0 calls), autograd, torch.optim.AdamW — timed beside this repo's CUDA path on the same
synthetic batch.  It is a baseline, not a target; the test asserts only that the hand-written path is not slower.
"""
import json

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
dev = "cuda"


def test_step_time_vs_torch_library_path(built_lib, monkeypatch):
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from oracle import clip_ref, ddpm_ref, step_ref, unet_ref
    from textboost_b200 import synthetic

    B, V = 8, 49408

    def sdpa_forward(self, x, ctx=None):  # diffusers AttnProcessor2_0: F.scaled_dot_product_attention
        ctx = x if ctx is None else ctx
        Bq, N, C = x.shape
        d = C // self.heads
        q = self.to_q(x).view(Bq, N, self.heads, d).transpose(1, 2)
        k = self.to_k(ctx).view(Bq, -1, self.heads, d).transpose(1, 2)
        v = self.to_v(ctx).view(Bq, -1, self.heads, d).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(Bq, N, C)
        return self.to_out[0](o)

    monkeypatch.setattr(unet_ref.Attention, "forward", sdpa_forward)
    torch.backends.cudnn.benchmark = True
    with torch.no_grad():
        unet = unet_ref.init_unet_(unet_ref.UNet2DConditionModelRef(unet_ref.UNetConfig.sd15()), 0).to(dev).half()
        unet.requires_grad_(False)
        ccfg = clip_ref.ClipTextConfig.clip_l()
        te0 = clip_ref.init_clip_(clip_ref.TextBoostModelRef(ccfg), 1)
        te0.set_null_embedding(torch.randn(77, ccfg.hidden_size, generator=torch.Generator().manual_seed(2)))
        import copy
        te = copy.deepcopy(te0)
        te0 = te0.to(dev).half().requires_grad_(False)   # train_textboost.py:939
        te.resize_token_embeddings(V + 1)
        te.get_input_embeddings().weight[V:] = te.get_input_embeddings().weight[1929:1930]
    te.requires_grad_(False)
    te.add_adapter(r=4)
    te = te.to(dev)                                       # fp32 master weights, autocast forward (:919-922)
    te.get_input_embeddings().weight.requires_grad_(True)
    opt = step_ref.make_optimizer(te)
    bt = synthetic.batch(B, 64, 42, V, dev)
    scale = 65536.0

    def library_step():
        noisy = ddpm_ref.add_noise(bt["latents"], bt["noise"], bt["timesteps"])
        with torch.autocast("cuda", dtype=torch.float16):
            ehs = te(bt["input_ids"])
        pred = unet(noisy.half(), bt["timesteps"], ehs.half())
        loss = F.mse_loss(pred.float(), bt["noise"].float(), reduction="none").mean()
        with torch.autocast("cuda", dtype=torch.float16):
            h = te(bt["prior_ids"])
        with torch.no_grad():
            h0 = te0(bt["prior_ids"])
        loss = loss + 0.1 * (1 - F.cosine_similarity(h.float(), h0.float(), dim=-1)).mean()
        (loss * scale).backward()
        emb = te.get_input_embeddings().weight
        emb.grad[:V] = 0
        lora = [p for n, p in te.named_parameters() if "lora_" in n]
        for p in lora:
            p.grad.div_(scale)
        emb.grad.div_(scale)
        torch.nn.utils.clip_grad_norm_(lora, 1.0)
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    def timed(fn, warm, n):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, out

    ms_lib, loss_lib = timed(library_step, 2, 3)
    assert torch.isfinite(loss_lib)
    del unet, te, te0, opt
    torch.cuda.empty_cache()

    tr = synthetic.build_trainer("sd15", dev, seed=42, n_added=1)
    args = (bt["latents"], bt["noise"], bt["timesteps"], bt["input_ids"], bt["prior_ids"])
    tr.step(*args)
    replay = tr.capture(*args, warmup=1)
    ms_ours, _ = timed(lambda: replay(*args), 3, 10)
    line = {"library_baseline": {"ms_per_step": ms_lib, "images_per_s": B / (ms_lib / 1e3),
                                 "what": "oracle modules in fp16 on the B200 via torch 2.11 (cuDNN conv, cuBLASLt, "
                                         "SDPA, autograd, torch AdamW), batch 8, KPL on"},
            "ours": {"ms_per_step": ms_ours, "images_per_s": B / (ms_ours / 1e3)},
            "speedup": ms_lib / ms_ours}
    print("LIBRARY_BASELINE " + json.dumps(line))
    assert ms_ours < ms_lib
