"""GPU probe: the image-fed training step (SD-1.5, batch 8, 512^2): pixels -> AutoencoderKL encoder -> latents -> step.

Three schedules over the same work: (a) encoder then step on one stream (what a plain loop does), (b) the encoder of
batch i+1 on a second stream while the captured step of batch i replays (the CLI draws batches one step ahead), and for
reference (c) the latents-fed step alone.  Pixels are resident on the device (the deferred-augmentation front end leaves
~1 ms of host work per batch, profiles/r02_frontend_probe.json).  Writes gpurun_out/<tag>_image_fed.json."""
import json
import sys

import torch

sys.path.insert(0, ".")
from textboost_b200 import synthetic, vae  # noqa: E402

dev = "cuda"
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
B, STEPS = 8, 12
tr = synthetic.build_trainer("sd15", dev, seed=42, n_added=1, kpl_weight=0.1)
cfg = vae.VAEConfig()
enc = vae.VAEEncoderEngine(cfg, synthetic.random_vae_sd(cfg, dev, 0), max_chunk=8)
bt = synthetic.batch(B, 64, 7, 49408, dev)
pixels = [torch.rand(B, 3, 512, 512, device=dev) * 2 - 1 for _ in range(2)]
gen = torch.Generator(device=dev).manual_seed(1)
args = [bt["latents"], bt["noise"], bt["timesteps"], bt["input_ids"], bt["prior_ids"]]
for _ in range(2):
    tr.step(*args)
replay = tr.capture(*args, warmup=1)
for _ in range(2):
    lat = enc.encode_latents(pixels[0], generator=gen)
    replay(lat, *args[1:])
torch.cuda.synchronize()


def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / STEPS


def serial():
    for i in range(STEPS):
        lat = enc.encode_latents(pixels[i & 1], generator=gen)
        replay(lat, *args[1:])


side = torch.cuda.Stream()


def overlapped():
    main = torch.cuda.current_stream()
    with torch.cuda.stream(side):
        side.wait_stream(main)
        nxt = enc.encode_latents(pixels[0], generator=gen)
        ready = torch.cuda.Event()
        ready.record(side)
    for i in range(STEPS):
        main.wait_event(ready)
        lat = nxt
        lat.record_stream(main)
        with torch.cuda.stream(side):  # batch i+1 is encoded while step i runs
            nxt = enc.encode_latents(pixels[(i + 1) & 1], generator=gen)
            ready = torch.cuda.Event()
            ready.record(side)
        replay(lat, *args[1:])
    main.wait_stream(side)


def step_only():
    for _ in range(STEPS):
        replay(*args)


out = {"batch": B, "steps": STEPS}
for name, fn in (("latents_fed_step", step_only), ("encoder_then_step", serial), ("encoder_overlapped", overlapped)):
    fn()
    ms = min(timed(fn) for _ in range(2))
    out[name] = {"ms_per_step": ms, "images_per_s": B * 1e3 / ms}
    print(name, out[name])
json.dump(out, open(f"gpurun_out/{tag}_image_fed.json", "w"), indent=1)
