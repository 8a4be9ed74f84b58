#!/bin/bash
# 2-GPU checks (gpurun --gpus 2): bench.py under torchrun exits cleanly; the CLI trains a tiny synthetic checkpoint on 2 ranks.
mkdir -p gpurun_out
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --hang-timeout 150 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench rc=$?"
grep '^{' gpurun_out/bench_2gpu.json | cut -c1-200
python - <<'PY'
import sys
sys.path.insert(0, ".")
from textboost_b200 import synthetic
synthetic.write_pretrained("/tmp/tiny_ckpt", "tiny", seed=2)
PY
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 train_textboost.py --pretrained_model_name_or_path /tmp/tiny_ckpt --output_dir /tmp/tiny_out --synthetic_data --resolution 128 --train_batch_size 2 --max_train_steps 12 --checkpointing_steps 6 --learning_rate 1e-3 --mixed_precision fp16 --log_every 4 > gpurun_out/cli_2gpu.log 2>&1; echo "cli rc=$?"
grep -E "step (4|8|12) " gpurun_out/cli_2gpu.log | tail -3; ls /tmp/tiny_out
