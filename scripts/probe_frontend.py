"""Time the image front end on one B200 (next-round measurement; not run this round, GPU budget spent): for a batch of
B items at 512^2 from 1024^2 sources — host time of the three dataset modes, GPU time of the deferred augmentation
(run_plan), the resize / crop / normalise tail and the VAE encoder — and check that the three modes produce the same
pixel_values bits.  Prints one JSON line and writes gpurun_out/frontend_probe.json.

    python scripts/probe_frontend.py [B=8] [steps=10]
"""
import json
import os
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_augment_golden as G  # noqa: E402  (deterministic synthetic images)
from textboost_b200 import augment, dataset, image_ops, synthetic, vae  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = "cuda"
out = {"batch": B, "steps": steps}
cfg = vae.VAEConfig()
enc = vae.VAEEncoderEngine(cfg, synthetic.random_vae_sd(cfg, dev, 0), max_chunk=8)
with tempfile.TemporaryDirectory() as d:
    for i in range(4):
        G.make_image((1024, 1024), i).save(os.path.join(d, f"{i}.jpg"), quality=92)
    concepts = [{"instance_data_dir": d, "instance_token": "<dog>"}]
    batches = {}
    for mode in ("host", "device_transforms", "device_augment"):
        pipe = augment.PairedAugmentation(hflip="inversion", inversion=True, p=0.8)
        ds = dataset.TextBoostDataset(concepts, synthetic.LiteralTokenizer(), size=512, augment_pipe=pipe,
                                      template="textboost", device_transforms=mode != "host",
                                      cache_decoded=mode != "host", device_augment=mode == "device_augment")
        G.seed_all(0)
        host_s, gpu_ms, px = 0.0, 0.0, None
        for s in range(steps + 2):
            t0 = time.perf_counter()
            batch = dataset.TextBoostDataset.collate_fn([ds[s * B + i] for i in range(B)], False)
            dt = time.perf_counter() - t0
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if "sources" in batch:
                px = image_ops.batch_to_pixel_values(batch["sources"], dev)
            else:
                px = batch["pixel_values"].to(dev)
            e1.record()
            torch.cuda.synchronize()
            if s >= 2:
                host_s += dt
                gpu_ms += e0.elapsed_time(e1)
            if s == 2:
                batches[mode] = px.clone()
        out[mode] = {"host_ms_per_batch": host_s / steps * 1e3, "gpu_ms_per_batch": gpu_ms / steps}
    out["same_bits"] = bool(torch.equal(batches["host"], batches["device_transforms"])
                            and torch.equal(batches["host"], batches["device_augment"]))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eps = torch.randn(B, 4, 64, 64, device=dev)
    for _ in range(2):
        enc.encode_latents(px, eps)
    e0.record()
    for _ in range(steps):
        enc.encode_latents(px, eps)
    e1.record()
    torch.cuda.synchronize()
    out["vae_encode_ms_per_batch"] = e0.elapsed_time(e1) / steps
print("FRONTEND_PROBE " + json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/frontend_probe.json", "w") as f:
    json.dump(out, f, indent=1)
