"""Run ONE eager TextBoost step (SD-1.5, B=8, KPL on) between cudaProfilerStart/Stop, for
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file ... python scripts/ncu_step.py
and for `ncu --set full -k regex:<kernel>` captures of single kernels.  Numbers printed under ncu are not bench values."""
import sys
import torch
sys.path.insert(0, ".")
from textboost_b200 import _cabi, synthetic  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
tr = synthetic.build_trainer("sd15", "cuda", seed=42, n_added=1)
bt = synthetic.batch(B, 64, 42, 49408, "cuda")
args = (bt["latents"], bt["noise"], bt["timesteps"], bt["input_ids"], bt["prior_ids"])
for _ in range(2):
    tr.step(*args)
torch.cuda.synchronize()
n0 = _cabi.launch_count
torch.cuda.profiler.start()
tr.step(*args)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launches in profiled step:", _cabi.launch_count - n0, "loss", tr.loss.item())
