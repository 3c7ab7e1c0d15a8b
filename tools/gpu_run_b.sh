#!/bin/bash
# round-2 GPU call B (1 GPU): re-check after the fixes of call A -- tests, bench line, ncu full capture of the TMA mover
mkdir -p gpurun_out
TAG=r02b
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -15 gpurun_out/${TAG}_pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -c 7000 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
echo "ref rc=$?"; tail -c 1500 gpurun_out/${TAG}_bench_reference.json
timeout 300 python tools/mover_sweep.py > gpurun_out/${TAG}_mover_sweep.jsonl 2>&1; tail -12 gpurun_out/${TAG}_mover_sweep.jsonl
TFW_MOVER=tma timeout 400 ncu --set full --clock-control none --import-source on -k regex:tfw_mover_tma1 -s 6 -c 2 -f -o gpurun_out/${TAG}_mover_tma1 \
    python bench.py --steps 2 --warmup 3 --latency-calls 50 --no-cpu-baseline --no-boundary --no-c3 --no-c4 --no-swap > gpurun_out/${TAG}_ncu_tma.log 2>&1
for T in 8 16 32 64; do TFW_PARK_THREADS=$T timeout 120 python -m pytest tests/test_gpu_isolation.py -q -k parked -s 2>&1 | grep -E "park|tfw\]" | tr '\n' ' '; echo " threads=$T"; done > gpurun_out/${TAG}_park_threads.txt 2>&1; cat gpurun_out/${TAG}_park_threads.txt
ls -la gpurun_out/ | tail -12
