#!/bin/bash
# tools/profile.sh <tag>  -- run on the GPU box (under gpurun): bench line, ncu launch list, one ncu --set full capture.
# Outputs land in gpurun_out/; copy the summaries into profiles/<tag>_* afterwards (tools/summarize_ncu.py).
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/${TAG}_clocks_before.csv
# 1) the bench line (never under a profiler)
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -1 gpurun_out/${TAG}_bench.json
# 2) every launch with its device time (same command line as the bench, fewer steps)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --latency-calls 50 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
# 3) the top kernel once, full set, with source
ncu --set full --clock-control none --import-source on -k regex:tfw_mover_ldg -s 6 -c 3 -f -o gpurun_out/${TAG}_mover \
    python bench.py --steps 2 --warmup 3 --latency-calls 50 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out/
