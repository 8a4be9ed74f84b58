"""GPU probe: text-encoder passes alone (CLIP-L, 8 prompts x 77 tokens, LoRA r=4 on q/k/v): captured-graph time of the
trainable forward (saving for backward), its backward, and the frozen forward, plus launches per pass."""
import sys

import torch

sys.path.insert(0, ".")
from textboost_b200 import _cabi as C  # noqa: E402
from textboost_b200 import synthetic  # noqa: E402

dev = "cuda"
tr = synthetic.build_trainer("sd15", dev, seed=42, n_added=1, kpl_weight=0.1)
te, te0 = tr.te, tr.te0
bt = synthetic.batch(8, 64, 7, 49408, dev)
ids = bt["input_ids"]
te.pack_lora()
d_out = torch.randn(8, 77, 768, device=dev)


def fwd():
    te.forward(ids, save_for_backward=True)


def fwd_bwd():
    te.forward(ids, save_for_backward=True)
    te.backward(d_out.clone())


def fwd0():
    te0.forward(ids)


def graph_time(fn, iters=20):
    fn()
    torch.cuda.synchronize()
    n0 = C.launch_count
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    n0 = C.launch_count
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    n = C.launch_count - n0
    te._ctx.clear()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, n


for name, fn in (("trainable forward", fwd), ("trainable forward + backward", fwd_bwd), ("frozen forward", fwd0)):
    ms, n = graph_time(fn)
    print(f"{name:32s} {ms * 1000:8.1f} us  {n:4d} launches")
