"""Cost of --unet_params_to_train crossattn_kv on the SD-1.5 batch-8 step (bf16 policy, the one the mode runs in):
the captured step with and without the UNet K/V adapter.

  python scripts/probe_cross_kv.py [--out gpurun_out/r02_probe_cross_kv.json]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from textboost_b200 import precision  # noqa: E402

precision.set_policy("bf16")
import torch  # noqa: E402

from textboost_b200 import _cabi, synthetic  # noqa: E402


def timed(replay, args, steps=20, warmup=5):
    for _ in range(warmup):
        replay(*args)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        replay(*args)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/r02_probe_cross_kv.json")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    res = {"policy": "bf16", "workload": "SD-1.5 512^2 bs=8, KPL on, rank-4 adapters"}
    for name, r in (("text_encoder_lora_only", 0), ("plus_unet_crossattn_kv", 4)):
        tr = synthetic.build_trainer("sd15", dev, seed=42, n_added=1, kpl_weight=0.1, unet_lora_r=r)
        bt = synthetic.batch(8, 64, 42, 49408, dev)
        args = (bt["latents"], bt["noise"], bt["timesteps"], bt["input_ids"], bt["prior_ids"])
        n0 = _cabi.launch_count
        tr.step(*args)
        torch.cuda.synchronize()
        launches = _cabi.launch_count - n0
        replay = tr.capture(*args, warmup=1)
        ms = timed(replay, args)
        res[name] = {"ms_per_step": ms, "images_per_s": 8 / ms * 1e3, "launches_per_step": launches,
                     "loss": float(tr.loss.item()),
                     "trainable_floats": int(tr.te.state.params.numel() + (tr.opt_unet.params.numel() if r else 0))}
        print(name, res[name], flush=True)
        del tr, replay
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
