#!/bin/bash
# One gpurun call at the end of a round: launch list + DRAM traffic of one eager step under ncu (summarised on the box:
# the raw logs stay small), the GPU test suite, smoke, and the bench line.  Usage: bash scripts/gpu_final_evidence.sh r02
tag=${1:-r02}
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file /tmp/${tag}_traffic.csv python scripts/ncu_step.py > gpurun_out/${tag}_ncu_step.log 2>&1
echo "ncu rc=$?"; tail -1 gpurun_out/${tag}_ncu_step.log
python scripts/traffic_from_ncu.py /tmp/${tag}_traffic.csv gpurun_out/${tag}_traffic.json && cat gpurun_out/${tag}_traffic.json | head -c 700; echo
grep -v "dram__bytes" /tmp/${tag}_traffic.csv > /tmp/${tag}_launches.csv
python scripts/summarize_launches.py /tmp/${tag}_launches.csv > gpurun_out/${tag}_launches_summary_final.txt 2>&1; head -12 gpurun_out/${tag}_launches_summary_final.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${tag}_pytest_gpu_final.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${tag}_smoke.log
timeout 600 python bench.py > gpurun_out/${tag}_bench_final.json 2> gpurun_out/${tag}_bench_final.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/${tag}_bench_final.json
