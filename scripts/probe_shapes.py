"""Per-entry-point and per-shape GPU time of one eager TextBoost step (SD-1.5, B=8, KPL on).

CUDA events bracket every C-ABI call.  The GPU is first parked behind a long sleep kernel so the host runs
ahead and the kernels execute back to back (event gaps then measure kernels, not Python launch latency).

  python scripts/probe_shapes.py [B] > gpurun_out/shapes.txt
"""
import sys
from collections import defaultdict

import torch

sys.path.insert(0, ".")
from textboost_b200 import _cabi, ops, synthetic  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
tr = synthetic.build_trainer("sd15", "cuda", seed=42, n_added=1)
bt = synthetic.batch(B, 64, 42, 49408, "cuda")
args = (bt["latents"], bt["noise"], bt["timesteps"], bt["input_ids"], bt["prior_ids"])
for _ in range(3):
    tr.step(*args)
torch.cuda.synchronize()

rec = []
orig_call = _cabi.call


def shape_sig(name, a):
    v = [x.value if hasattr(x, "value") else x for x in a]
    if name == "tb_gemm_f16":
        return f"M={v[6]} N={v[7]} K={v[8]}", 2.0 * v[6] * v[7] * v[8]
    if name == "tb_conv3x3_f16":
        Bb, H, W, Ci, Co = v[3:8]
        return f"B={Bb} H={H} Cin={Ci} Cout={Co} (M={Bb*H*W} N={Co} K={9*Ci})", 2.0 * Bb * H * W * Co * 9 * Ci
    if name == "tb_attn_fwd_f16":
        Bb, h, Nq, Nk, d = v[9:14]
        return f"B={Bb} h={h} Nq={Nq} Nk={Nk} d={d}", 4.0 * Bb * h * Nq * Nk * d
    if name == "tb_attn_bwd_f16":
        Bb, h, Nq, Nk, d = v[18:23]
        dq = v[12] is not None and v[12] != 0
        return f"B={Bb} h={h} Nq={Nq} Nk={Nk} d={d} dq={int(bool(dq))}", (8.0 if dq else 6.0) * Bb * h * Nq * Nk * d
    if name in ("tb_groupnorm_fwd_f16", "tb_groupnorm_bwd_f16"):
        i = 5 if name.endswith("fwd_f16") else 8
        Bb, HW, Cc = v[i:i + 3]
        return f"B={Bb} HW={HW} C={Cc}", float(Bb * HW * Cc)
    return "", 0.0


def timed_call(name, *a):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    orig_call(name, *a)
    e1.record()
    sig, work = shape_sig(name, a)
    rec.append((name, sig, work, e0, e1))


_cabi.call = timed_call
ops.C.call = timed_call
torch.cuda._sleep(int(0.4 * 1.9e9))
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
tr.step(*args)
t1.record()
torch.cuda.synchronize()
_cabi.call = orig_call
total = t0.elapsed_time(t1)

by_name = defaultdict(lambda: [0, 0.0, 0.0])
by_shape = defaultdict(lambda: [0, 0.0, 0.0])
for name, sig, work, e0, e1 in rec:
    ms = e0.elapsed_time(e1)
    for d, k in ((by_name, name), (by_shape, (name, sig))):
        d[k][0] += 1
        d[k][1] += ms
        d[k][2] += work
acc = sum(v[1] for v in by_name.values())
print(f"step (events, host ahead): {total:.3f} ms; sum of bracketed calls {acc:.3f} ms; {len(rec)} calls")
print(f"\n{'entry point':28s} {'calls':>6s} {'ms':>9s} {'share':>7s} {'TFLOP/s':>9s}")
for k, (n, ms, w) in sorted(by_name.items(), key=lambda kv: -kv[1][1]):
    tf = w / ms / 1e9 if w and k not in ("tb_groupnorm_fwd_f16", "tb_groupnorm_bwd_f16") else 0
    print(f"{k:28s} {n:6d} {ms:9.3f} {100*ms/acc:6.1f}% {tf:9.1f}")
print(f"\n{'entry point / shape':86s} {'calls':>5s} {'ms':>8s} {'us/call':>8s} {'TFLOP/s|Gelem/s':>10s}")
for (name, sig), (n, ms, w) in sorted(by_shape.items(), key=lambda kv: -kv[1][1]):
    if not sig:
        continue
    rate = w / ms / 1e9 if name not in ("tb_groupnorm_fwd_f16", "tb_groupnorm_bwd_f16") else w / ms / 1e6
    print(f"{(name[3:] + ' ' + sig):86s} {n:5d} {ms:8.3f} {1e3*ms/n:8.1f} {rate:10.1f}")
