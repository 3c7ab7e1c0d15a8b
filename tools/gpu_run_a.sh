#!/bin/bash
# round-2 GPU call (1 GPU): environment facts, the GPU test-suite, the process-boundary bench, the bench line of both arms,
# refreshed kernel evidence (copy lab incl. the TMA one-shot shape, mover sweep, provider round trips), ncu launch list and
# one full capture of the roofline kernel
mkdir -p gpurun_out
TAG=r02
{ df -h /dev/shm /tmp; nproc; free -g; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv; sysctl net.core.rmem_max net.core.wmem_max; cat /sys/kernel/mm/transparent_hugepage/enabled; } > gpurun_out/${TAG}_env.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -15 gpurun_out/${TAG}_pytest_gpu.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/${TAG}_clocks_before.csv
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -c 7000 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
echo "ref rc=$?"; tail -c 1500 gpurun_out/${TAG}_bench_reference.json
timeout 300 tools/copy_lab > gpurun_out/${TAG}_copy_lab.jsonl 2>&1; grep -E "oneshot|cudaMemcpy" gpurun_out/${TAG}_copy_lab.jsonl | head -12
timeout 300 python tools/mover_sweep.py > gpurun_out/${TAG}_mover_sweep.jsonl 2>&1; tail -16 gpurun_out/${TAG}_mover_sweep.jsonl
timeout 200 python tools/cpu_baselines.py > gpurun_out/${TAG}_cpu_baselines.jsonl 2>&1; tail -4 gpurun_out/${TAG}_cpu_baselines.jsonl
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --latency-calls 50 --no-cpu-baseline --no-boundary --no-c3 --no-c4 > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tfw_mover_ldg -s 6 -c 2 -f -o gpurun_out/${TAG}_mover_ldg \
    python bench.py --steps 2 --warmup 3 --latency-calls 50 --no-cpu-baseline --no-boundary --no-c3 --no-c4 --no-swap > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out/ | tail -20
