"""GPU probe: where the main stream's time goes inside one TextBoost step (SD-1.5, batch 8, KPL on).

CUDA events at module boundaries of an eager step with the host running ahead (the GPU is parked behind a sleep kernel),
so the intervals are GPU time on the main stream including whatever the side stream (text-encoder passes) takes from
it.  Also: the captured step with and without the knowledge-preservation branch.  Writes gpurun_out/<tag>_timeline.txt.
"""
import collections
import sys

import torch

sys.path.insert(0, ".")
from textboost_b200 import synthetic, unet as U  # noqa: E402

dev = "cuda"
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
out = open(f"gpurun_out/{tag}_timeline.txt", "w")


def say(*a):
    s = " ".join(str(x) for x in a)
    print(s)
    out.write(s + "\n")


def graph_ms(tr, bt, iters=10):
    args = (bt["latents"], bt["noise"], bt["timesteps"], bt["input_ids"], bt["prior_ids"])
    for _ in range(2):
        tr.step(*args)
    replay = tr.capture(*args, warmup=1)
    for _ in range(3):
        replay(*args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        replay(*args)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    B = 8
    tr = synthetic.build_trainer("sd15", dev, seed=42, n_added=1, kpl_weight=0.1)
    bt = synthetic.batch(B, 64, 7, 49408, dev)
    args = (bt["latents"], bt["noise"], bt["timesteps"], bt["input_ids"], bt["prior_ids"])
    for _ in range(2):
        tr.step(*args)
    torch.cuda.synchronize()

    marks = []  # (name, start event, end event)

    def wrap(obj, meth, name):
        fn = getattr(obj, meth)

        def inner(*a, **kw):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = fn(*a, **kw)
            e.record()
            marks.append((name, s, e))
            return r
        setattr(obj, meth, inner)
        return fn

    un = tr.unet
    saved = []

    def wrap_mod(m, name):
        for meth in ("forward", "backward"):
            saved.append((m, meth, wrap(m, meth, f"{name}.{meth[:3]}")))

    for i, blk in enumerate(un.down):
        for j, r in enumerate(blk["res"]):
            wrap_mod(r, f"down{i}.res{j}")
        for j, a in enumerate(blk["attn"]):
            wrap_mod(a, f"down{i}.attn{j}")
        if blk["down"] is not None:
            wrap_mod(blk["down"], f"down{i}.down")
    wrap_mod(un.mid_res[0], "mid.res0")
    wrap_mod(un.mid_attn, "mid.attn")
    wrap_mod(un.mid_res[1], "mid.res1")
    for i, blk in enumerate(un.up):
        for j, r in enumerate(blk["res"]):
            wrap_mod(r, f"up{i}.res{j}")
        for j, a in enumerate(blk["attn"]):
            wrap_mod(a, f"up{i}.attn{j}")
        if blk["up"] is not None:
            wrap_mod(blk["up"], f"up{i}.up")
    saved.append((un, "forward", wrap(un, "forward", "UNET.fwd")))
    saved.append((un, "backward", wrap(un, "backward", "UNET.bwd")))
    saved.append((tr.te, "backward", wrap(tr.te, "backward", "CLIP.bwd")))
    saved.append((tr.te, "forward", wrap(tr.te, "forward", "CLIP.fwd")))
    saved.append((tr.opt, "step", wrap(tr.opt, "step", "OPT.step")))

    torch.cuda.synchronize()
    torch.cuda._sleep(int(0.3 * 1.9e9))
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    tr.step(*args)
    s1.record()
    torch.cuda.synchronize()
    say(f"eager step with {len(marks)} bracketed modules: {s0.elapsed_time(s1):.3f} ms")
    say(f"{'module':24s} {'start':>8s} {'ms':>8s}")
    agg = collections.OrderedDict()
    for name, s, e in marks:
        ms = s.elapsed_time(e)
        say(f"{name:24s} {s0.elapsed_time(s):8.3f} {ms:8.3f}")
        agg[name] = ms
    for m, meth, fn in saved:
        setattr(m, meth, fn)

    # per level sums
    lev = collections.defaultdict(float)
    for name, ms in agg.items():
        if name.startswith(("UNET", "CLIP", "OPT")):
            continue
        blk, rest = name.split(".", 1)
        mod = rest.split(".")[0].rstrip("0123456789")
        lev[(blk, mod, rest.split(".")[-1])] += ms
    say("\nper (block, module kind, direction):")
    for k in sorted(lev):
        say(f"  {k[0]:6s} {k[1]:5s} {k[2]:4s} {lev[k]:8.3f} ms")

    say(f"\ncaptured step, KPL on : {graph_ms(tr, bt):.3f} ms")
    del tr
    torch.cuda.empty_cache()
    tr0 = synthetic.build_trainer("sd15", dev, seed=42, n_added=1, kpl_weight=0.0)
    bt0 = dict(bt)
    bt0["prior_ids"] = None
    say(f"captured step, KPL off: {graph_ms(tr0, bt0):.3f} ms")


main()
out.close()
