#!/bin/bash
# First gpurun call of the next round: everything that was built after round 1's GPU budget ran out, in one call.
#   gpurun --timeout 900 -- 'bash scripts/gpu_round2_first_call.sh'
# 1. the whole GPU suite with xfail details (-rxX lists which of the CPU-pinned, xfail-marked tests XPASSed / XFAILed)
# 2. the late-built probes: sampler (images/s, ms per denoising step), image front end (host / GPU ms per batch of the three
#    dataset modes, bit-equality of their pixel_values, VAE encode)
# 3. the bench line, for continuity
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -rxX > gpurun_out/pytest_gpu_r02_first.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu_r02_first.log
timeout 240 python scripts/probe_frontend.py 8 10 > gpurun_out/frontend_probe.log 2>&1; echo "frontend rc=$?"; tail -2 gpurun_out/frontend_probe.log
timeout 300 python scripts/probe_sampler.py 4 25 sd15 > gpurun_out/sampler_probe.log 2>&1; echo "sampler rc=$?"; tail -2 gpurun_out/sampler_probe.log
timeout 400 python bench.py > gpurun_out/bench_r02_first.json 2> gpurun_out/bench_r02_first.err; echo "bench rc=$?"; cat gpurun_out/bench_r02_first.json
