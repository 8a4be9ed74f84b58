"""GPU probe: UNetEngine forward / backward(d ehs) vs the PyTorch oracle (fp32 on the same GPU)."""
import sys
import time
import torch
sys.path.insert(0, ".")
from oracle import unet_ref  # noqa: E402
from textboost_b200 import unet as U  # noqa: E402

torch.manual_seed(0)
dev = "cuda"
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def run(name, rcfg, cfg, B, HW, L, check=True, timing=False):
    ref = unet_ref.UNet2DConditionModelRef(rcfg)
    unet_ref.init_unet_(ref, seed=1)
    ref = ref.to(dev)
    eng = U.UNetEngine(cfg, {k: v for k, v in ref.state_dict().items()})
    g = torch.Generator(device="cpu").manual_seed(2)
    x = torch.randn(B, 4, HW, HW, generator=g).to(dev)
    t = torch.randint(0, 1000, (B,), generator=g).to(dev)
    ehs = torch.randn(B, L, rcfg.cross_attention_dim, generator=g).to(dev)
    dout = torch.randn(B, 4, HW, HW, generator=g).to(dev) * 0.1
    x16, e16, d16 = x.half(), ehs.half(), dout.half()
    out = eng.forward(x16, t, e16)
    torch.cuda.synchronize()
    d_ehs = eng.backward(d16)
    torch.cuda.synchronize()
    print(f"[{name}] out finite={torch.isfinite(out).all().item()} std={out.float().std().item():.4f} "
          f"d_ehs finite={torch.isfinite(d_ehs).all().item()} std={d_ehs.std().item():.3e}", flush=True)
    if check:
        for p in ref.parameters():
            p.requires_grad_(False)
        # reference on the fp16-rounded inputs and weights, computed in fp32
        ref16 = ref  # weights already representable? round them like the engine does
        with torch.no_grad():
            for p in ref16.parameters():
                p.copy_(p.half().float())
        er = e16.float().requires_grad_(True)
        oref = ref16(x16.float(), t, er)
        oref.backward(d16.float())
        eo = ((out.float() - oref).abs().max() / oref.abs().max()).item()
        eg = ((d_ehs - er.grad).abs().max() / er.grad.abs().max()).item()
        cos = torch.nn.functional.cosine_similarity(d_ehs.flatten(), er.grad.flatten(), dim=0).item()
        print(f"[{name}] out relerr(max/max)={eo:.3e}  d_ehs relerr={eg:.3e} cos={cos:.6f}", flush=True)
        # fp16 torch baseline error envelope (what the reference's own fp16 path would give)
        ref_h = ref16.half()
        with torch.no_grad():
            oh = ref_h(x16, t, e16)
        print(f"[{name}] torch-fp16 forward relerr vs fp32 = {((oh.float() - oref).abs().max() / oref.abs().max()).item():.3e}", flush=True)
        ref_h.float()
    if timing:
        for _ in range(2):
            eng.forward(x16, t, e16)
            eng.backward(d16)
        torch.cuda.synchronize()
        e0, e1, e2 = torch.cuda.Event(True), torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(3):
            eng.forward(x16, t, e16)
        e1.record()
        for _ in range(3):
            eng.forward(x16, t, e16)
            eng.backward(d16)
        e2.record()
        torch.cuda.synchronize()
        tf = e0.elapsed_time(e1) / 3
        tfb = e1.elapsed_time(e2) / 3
        print(f"[{name}] B={B} fwd {tf:.2f} ms  fwd+bwd {tfb:.2f} ms  -> {B / tfb * 1e3:.1f} img/s (UNet only)", flush=True)
    del eng, ref
    torch.cuda.empty_cache()


tiny = unet_ref.UNetConfig.tiny()
run("tiny", tiny, U.UNetConfig(block_out_channels=tiny.block_out_channels,
                               attention_head_dim=tiny.attention_head_dim,
                               cross_attention_dim=tiny.cross_attention_dim, sample_size=16), 2, 16, 77)
run("sd15-b2", unet_ref.UNetConfig.sd15(), U.UNetConfig.sd15(), 2, 64, 77)
run("sd15-b8", unet_ref.UNetConfig.sd15(), U.UNetConfig.sd15(), 8, 64, 77, check=False, timing=True)
