"""Library baselines on the same B200: what torch 2.11's Blackwell library
kernels -- cuBLASLt linears, cuDNN convolutions (channels_last), cuDNN / flash SDPA, native GroupNorm / LayerNorm --
take for the shapes of the SD-1.5 step, beside this repo's kernels, and the whole reference step (oracle modules, fp16,
channels_last, autograd, capturable AdamW) replayed as ONE CUDA graph -- the fair end-to-end library number.

Every kernel pair is timed the same way: 20 back-to-back calls captured in a CUDA graph (no host launch overhead on
either side), 3 warm-up replays, 5 timed replays, CUDA events.  Writes gpurun_out/r2_library_baselines.json.

    python scripts/library_baselines.py [--no-step]
"""
import json
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from textboost_b200 import ops  # noqa: E402

dev = "cuda"
F16 = torch.float16
torch.backends.cudnn.benchmark = True
REP = 20


def graph_time(fn):
    """mean ms per call of fn, REP calls per graph replay."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(REP):
            fn()
    for _ in range(3):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * REP)


def rnd(*shape, s=1.0):
    return (torch.randn(*shape, device=dev) * s).to(F16)


rows = []


def add(kind, shape, ours_ms, lib_ms, work, unit):
    rows.append({"kind": kind, "shape": shape, "ours_us": round(ours_ms * 1e3, 2), "library_us": round(lib_ms * 1e3, 2),
                 "ours_over_library": round(ours_ms / lib_ms, 3),
                 "ours_rate": round(work / ours_ms / 1e9, 1), "library_rate": round(work / lib_ms / 1e9, 1), "unit": unit})
    print(f"{kind:10s} {shape:44s} ours {ours_ms * 1e3:8.1f} us  library {lib_ms * 1e3:8.1f} us  ratio {ours_ms / lib_ms:5.2f}",
          flush=True)


def gemm_cases():
    # (M, N, K, residual): linears / 1x1 convs of the UNet at B = 8 and the text encoder at 616 rows
    for (M, N, K, res) in [(32768, 320, 320, True), (8192, 640, 640, True), (2048, 1280, 1280, True),
                           (32768, 2560, 320, False), (32768, 320, 1280, True), (8192, 5120, 640, False),
                           (2048, 10240, 1280, False), (2048, 1280, 5120, True), (32768, 960, 320, False),
                           (616, 768, 768, False), (616, 3072, 768, False), (616, 768, 3072, False),
                           (616, 2304, 784, False)]:
        a, w, b = rnd(M, K), rnd(N, K, s=K ** -0.5), rnd(N, s=0.1)
        r = rnd(M, N) if res else None
        ours = graph_time(lambda: ops.gemm(a, w, bias=b, residual=r))
        if res:
            lib = graph_time(lambda: torch.add(F.linear(a, w, b), r))
        else:
            lib = graph_time(lambda: F.linear(a, w, b))
        add("gemm", f"[{M},{K}]x[{N},{K}]^T+bias" + ("+res" if res else ""), ours, lib, 2.0 * M * N * K / 1e3, "TFLOP/s")


def conv_cases():
    for (B, H, Cin, Cout) in [(8, 64, 320, 320), (8, 32, 640, 640), (8, 16, 1280, 1280), (8, 8, 1280, 1280),
                              (8, 64, 640, 320), (8, 32, 1280, 640), (8, 16, 2560, 1280)]:
        x = rnd(B, H, H, Cin)
        w4 = rnd(Cout, Cin, 3, 3, s=(9 * Cin) ** -0.5)
        bias = rnd(Cout, s=0.1)
        wk = w4.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous()
        ours = graph_time(lambda: ops.conv3x3(x, wk, bias=bias))
        xc = x.permute(0, 3, 1, 2)  # NCHW view of NHWC storage == channels_last
        wc = w4.contiguous(memory_format=torch.channels_last)
        lib = graph_time(lambda: F.conv2d(xc, wc, bias, padding=1))
        add("conv3x3", f"B={B} {H}x{H} {Cin}->{Cout}", ours, lib, 2.0 * B * H * H * Cout * 9 * Cin / 1e3, "TFLOP/s")


def attn_cases():
    for (B, Hh, Nq, Nk, d) in [(8, 8, 4096, 4096, 40), (8, 8, 1024, 1024, 80), (8, 8, 256, 256, 160),
                               (8, 8, 4096, 77, 40), (8, 8, 1024, 77, 80), (8, 12, 77, 77, 64)]:
        C = Hh * d
        q, k, v, do = rnd(B, Nq, C), rnd(B, Nk, C), rnd(B, Nk, C), rnd(B, Nq, C)
        causal = Nq == 77
        o, lse = ops.attn_fwd(q, k, v, Hh, causal=causal)
        ours_f = graph_time(lambda: ops.attn_fwd(q, k, v, Hh, causal=causal))
        ours_b = graph_time(lambda: ops.attn_bwd(q, k, v, o, do, lse, Hh, causal=causal))

        def hd(t):
            return t.view(B, -1, Hh, d).transpose(1, 2)
        qh, kh, vh = (hd(t).detach().requires_grad_(True) for t in (q, k, v))
        lib_f = graph_time(lambda: F.scaled_dot_product_attention(qh.detach(), kh.detach(), vh.detach(), is_causal=causal))
        doh = hd(do)

        def fwd_bwd():  # (the backward of a graph built outside the capture would run on the legacy stream)
            oh = F.scaled_dot_product_attention(qh, kh, vh, is_causal=causal)
            return torch.autograd.grad(oh, (qh, kh, vh), doh)
        lib_b = graph_time(fwd_bwd) - lib_f
        fl = 4.0 * B * Hh * Nq * Nk * d / 1e3
        add("attn fwd", f"B={B} h={Hh} Nq={Nq} Nk={Nk} d={d}" + (" causal" if causal else ""), ours_f, lib_f, fl, "TFLOP/s")
        add("attn bwd", f"B={B} h={Hh} Nq={Nq} Nk={Nk} d={d}" + (" causal" if causal else "") + " (lib: fwd+bwd - fwd)",
            ours_b, lib_b, 2 * fl, "TFLOP/s")


def norm_cases():
    for (B, HW, C) in [(8, 4096, 320), (8, 1024, 640), (8, 256, 1280)]:
        x = rnd(B, HW, C)
        g, b = (1 + 0.1 * torch.randn(C, device=dev)).to(F16), rnd(C, s=0.1)
        ours = graph_time(lambda: ops.groupnorm(x, g, b, 32, 1e-5, True))
        xc = x.view(B, int(HW ** 0.5), int(HW ** 0.5), C).permute(0, 3, 1, 2)
        lib = graph_time(lambda: F.silu(F.group_norm(xc, 32, g, b, 1e-5)))
        add("gn+silu", f"[{B},{HW},{C}]", ours, lib, 4.0 * x.numel() / 1e3, "TB/s (2R... algorithmic 1R+1W)")
        x2 = x.view(B * HW, C)
        ours = graph_time(lambda: ops.layernorm(x2, g, b))
        lib = graph_time(lambda: F.layer_norm(x2, (C,), g, b, 1e-5))
        add("layernorm", f"[{B * HW},{C}]", ours, lib, 4.0 * x.numel() / 1e3, "TB/s")


def library_step_graphed():
    """The reference step through the oracle modules: fp16 UNet in channels_last, SDPA attention, autocast text
    encoder, scaled backward, capturable AdamW -- captured as one CUDA graph."""
    import copy
    from oracle import clip_ref, ddpm_ref, step_ref, unet_ref
    from textboost_b200 import synthetic
    B, V = 8, 49408

    def sdpa_forward(self, x, ctx=None):  # diffusers AttnProcessor2_0
        ctx = x if ctx is None else ctx
        Bq, N, C = x.shape
        d = C // self.heads
        q = self.to_q(x).view(Bq, N, self.heads, d).transpose(1, 2)
        k = self.to_k(ctx).view(Bq, -1, self.heads, d).transpose(1, 2)
        v = self.to_v(ctx).view(Bq, -1, self.heads, d).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(Bq, N, C)
        return self.to_out[0](o)

    unet_ref.Attention.forward = sdpa_forward

    def capturable_forward(self, input_ids):
        """TextBoostModel.forward without its data-dependent branch (`if null_pos.any()`, text_encoder.py:71, forces a
        host sync and cannot be captured): same result through torch.where -- the library step gets the benefit of a
        CUDA graph that the reference itself could not have."""
        out = self.text_model(input_ids)
        null_pos = (input_ids[:, 1] == self.EOS_ID)[:, None, None]
        out = torch.where(null_pos, self.null_embedding.to(out.dtype).unsqueeze(0), out)
        if self._use_fixed_special_embedding:
            out = torch.cat([self.null_embedding[0].to(out.dtype).expand(out.shape[0], 1, -1), out[:, 1:]], 1)
        return out

    clip_ref.TextBoostModelRef.forward = capturable_forward
    with torch.no_grad():
        unet = unet_ref.init_unet_(unet_ref.UNet2DConditionModelRef(unet_ref.UNetConfig.sd15()), 0).to(dev).half()
        unet = unet.to(memory_format=torch.channels_last).requires_grad_(False)
        ccfg = clip_ref.ClipTextConfig.clip_l()
        te0 = clip_ref.init_clip_(clip_ref.TextBoostModelRef(ccfg), 1)
        te0.set_null_embedding(torch.randn(77, ccfg.hidden_size, generator=torch.Generator().manual_seed(2)))
        te = copy.deepcopy(te0)
        te0 = te0.to(dev).half().requires_grad_(False)
        te.resize_token_embeddings(V + 1)
        te.get_input_embeddings().weight[V:] = te.get_input_embeddings().weight[1929:1930]
    te.requires_grad_(False)
    te.add_adapter(r=4)
    te = te.to(dev)
    te.get_input_embeddings().weight.requires_grad_(True)
    emb = te.get_input_embeddings().weight
    lora = [p for n, p in te.named_parameters() if "lora_" in n]
    opt = torch.optim.AdamW([{"params": [emb], "lr": 1e-3}, {"params": lora}], lr=5e-5, weight_decay=1e-2,
                            capturable=True)
    bt = synthetic.batch(B, 64, 42, V, dev)
    scale = 65536.0
    acp = ddpm_ref.alphas_cumprod().to(dev)  # (built on the host by default: a copy inside the capture otherwise)

    def step():
        noisy = ddpm_ref.add_noise(bt["latents"], bt["noise"], bt["timesteps"], acp).contiguous(memory_format=torch.channels_last)
        with torch.autocast("cuda", dtype=F16):
            ehs = te(bt["input_ids"])
        pred = unet(noisy.half(), bt["timesteps"], ehs.half())
        loss = F.mse_loss(pred.float(), bt["noise"].float(), reduction="none").mean()
        with torch.autocast("cuda", dtype=F16):
            h = te(bt["prior_ids"])
        with torch.no_grad():
            h0 = te0(bt["prior_ids"])
        loss = loss + 0.1 * (1 - F.cosine_similarity(h.float(), h0.float(), dim=-1)).mean()
        (loss * scale).backward()
        emb.grad[:V] = 0
        for p_ in lora:
            p_.grad.div_(scale)
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
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    eager = timed(step, 3, 5)
    out = {"eager_channels_last_ms": eager}
    try:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step()
        out["graphed_channels_last_ms"] = timed(g.replay, 3, 10)
    except Exception as e:  # noqa: BLE001
        import traceback
        out["graph_capture_error"] = repr(e)[:300]
        out["graph_capture_where"] = [l.strip() for l in traceback.format_exc().splitlines() if "File" in l][-6:]
    del unet, te, te0, opt
    torch.cuda.empty_cache()
    tr = synthetic.build_trainer("sd15", dev, seed=42, n_added=1)
    args = (bt["latents"], bt["noise"], bt["timesteps"], bt["input_ids"], bt["prior_ids"])
    tr.step(*args)
    replay = tr.capture(*args, warmup=1)
    out["ours_graphed_ms"] = timed(lambda: replay(*args), 3, 10)
    print("library step:", out, flush=True)
    return out


if __name__ == "__main__":
    res = {}
    if "--step-only" in sys.argv:
        print(json.dumps(library_step_graphed()))
        sys.exit(0)
    for case in (gemm_cases, conv_cases, attn_cases, norm_cases):
        try:
            case()
        except Exception as e:  # noqa: BLE001
            print(f"{case.__name__} failed: {e!r}"[:500], flush=True)
    res["kernels"] = rows
    if "--no-step" not in sys.argv:
        res["step"] = library_step_graphed()
    res["how"] = __doc__.split("\n\n")[1]
    with open("gpurun_out/r2_library_baselines.json", "w") as f:
        json.dump(res, f, indent=1)
