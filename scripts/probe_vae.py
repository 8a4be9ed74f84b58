"""Time the AutoencoderKL encoder engine (SD VAE config, random weights) on one B200: images/s at 512^2, CUDA events
on the launching stream, after warm-up; prints one JSON line and writes gpurun_out/vae_probe.json.
FLOPs per image: 1.117 TFLOP (558.3 GMAC)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from textboost_b200 import _cabi, synthetic, vae  # noqa: E402

dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
S = int(sys.argv[2]) if len(sys.argv) > 2 else 512
cfg = vae.VAEConfig()
eng = vae.VAEEncoderEngine(cfg, synthetic.random_vae_sd(cfg, dev, 0))
px = torch.rand(B, 3, S, S, device=dev) * 2 - 1
eps = torch.randn(B, 4, S // 8, S // 8, device=dev)
out = {"B": B, "size": S}
for chunk in (4, 8):
    eng.max_chunk = chunk
    for _ in range(2):
        lat = eng.encode_latents(px, eps)
    torch.cuda.synchronize()
    n0 = _cabi.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        lat = eng.encode_latents(px, eps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    out[f"chunk{chunk}"] = {"ms": ms, "images_per_s": B / ms * 1e3, "tflops": 1.117 * (S / 512) ** 2 * B / ms * 1e3 / 1e3,
                            "launches": (_cabi.launch_count - n0) // 5}
out["finite"] = bool(torch.isfinite(lat).all())
out["peak_mem_gb"] = torch.cuda.max_memory_allocated() / 2 ** 30
print("VAE_PROBE " + json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/vae_probe.json", "w") as f:
    json.dump(out, f, indent=1)
