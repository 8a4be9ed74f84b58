"""configs[4] at its real size (SD-2.x UNet + OpenCLIP-H, 768^2 images = 96x96 latents, batch 4): ms per captured step."""
import sys
import torch
sys.path.insert(0, ".")
from textboost_b200 import synthetic  # noqa: E402

tr = synthetic.build_trainer("sd21", "cuda", seed=7, n_added=1, prediction_type="v_prediction")
bt = synthetic.batch(4, 96, 3, 49408, "cuda")
args = (bt["latents"], bt["noise"], bt["timesteps"], bt["input_ids"], bt["prior_ids"])
tr.step(*args)
replay = tr.capture(*args, warmup=1)
for _ in range(3):
    replay(*args)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    replay(*args)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"config 5: {ms:.2f} ms / step, {4 / (ms / 1e3):.1f} images/s, {4.998 * 4 / (ms / 1e3):.0f} TFLOP/s algorithmic "
      f"(4.998 TFLOP per image, SURVEY.md 8d); loss {tr.loss.item():.4f}; "
      f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
