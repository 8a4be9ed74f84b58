"""Time the validation / inference sampler on one B200 (next-round measurement; not run this round, GPU budget spent):
SD-1.5-shaped random checkpoint written to a temp dir, `pipeline(prompt, num_images_per_prompt=N, num_inference_steps=S)`
with CUDA events around the denoising loop and the decode; prints one JSON line and writes gpurun_out/sampler_probe.json.

    python scripts/probe_sampler.py [N=4] [steps=25] [model=sd15]
Algorithmic FLOPs per image: steps x 2 x 0.8033 TFLOP (UNet forward, CFG doubled batch) + 2.51 TFLOP (VAE decoder)."""
import json
import os
import sys
import tempfile
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from textboost_b200 import _cabi, synthetic  # noqa: E402
from textboost_b200.pipeline import DPMSolverMultistepScheduler, StableDiffusionPipeline  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4
S = int(sys.argv[2]) if len(sys.argv) > 2 else 25
model = sys.argv[3] if len(sys.argv) > 3 else "sd15"
dev = "cuda"
out = {"images": N, "steps": S, "model": model}
with tempfile.TemporaryDirectory() as d:
    t0 = time.perf_counter()
    synthetic.write_pretrained(d, model, seed=0, vae_channels=(128, 256, 512, 512))
    pipe = StableDiffusionPipeline.from_pretrained(d, safety_checker=None)
    pipe.scheduler = DPMSolverMultistepScheduler.from_config(pipe.scheduler.config)
    pipe = pipe.to(dev)
    out["setup_s"] = time.perf_counter() - t0
for graph in (True, False):
    pipe.use_cuda_graph = graph
    pipe("a photo of a dog", num_images_per_prompt=N, num_inference_steps=3, output_type="pt")  # warm-up
    torch.cuda.synchronize()
    n0 = _cabi.launch_count
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    cond, uncond = pipe.encode_prompt("a photo of a dog", dev, N)
    x = pipe.prepare_latents(N, 512, 512, dev, generator=torch.Generator(device=dev).manual_seed(0))
    e[0].record()
    x = pipe.denoise(x, cond, uncond, S, 7.5)
    e[1].record()
    u8 = pipe.vae.decoder_engine.decode_u8(x)
    e[2].record()
    torch.cuda.synchronize()
    loop_ms, dec_ms = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
    tflop = N * (S * 2 * 0.8033 + 2.51)
    out["graph" if graph else "eager"] = {
        "denoise_ms": loop_ms, "ms_per_step": loop_ms / S, "decode_ms": dec_ms,
        "images_per_s": N / (loop_ms + dec_ms) * 1e3, "tflops": tflop / (loop_ms + dec_ms) * 1e3,
        "launches": _cabi.launch_count - n0, "finite": bool(torch.isfinite(x).all()), "u8_shape": list(u8.shape)}
print("SAMPLER_PROBE " + json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/sampler_probe.json", "w") as f:
    json.dump(out, f, indent=1)
