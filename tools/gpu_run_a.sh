#!/bin/bash
# round-2 GPU call A: environment facts, the GPU test-suite, the process-boundary bench
mkdir -p gpurun_out
{ df -h /dev/shm /tmp; nproc; free -g; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv; sysctl net.core.rmem_max net.core.wmem_max; } > gpurun_out/r02_env.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r02_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log
tail -5 gpurun_out/r02_pytest_gpu.log
timeout 900 python tools/boundary_bench.py > gpurun_out/r02_boundary.json 2> gpurun_out/r02_boundary.err
echo "boundary rc=$?"
cat gpurun_out/r02_boundary.json
tail -5 gpurun_out/r02_boundary.err
