#!/bin/bash
# round-2 last multi-GPU call (gpurun --gpus 8, charged 8x): the fixed-frames engine on the 1 TiB vGPU -- parity test, the
# sweep in two prefetch depths, then the 8-GPU bench line that embeds it next to the LRU engine
mkdir -p gpurun_out
TAG=r02g
timeout 200 python -m pytest tests/test_gpu_vram.py -q --timeout 150 -k "fixed or pipelined" > gpurun_out/${TAG}_pytest_vram_8gpu.log 2>&1; echo "vram rc=$?" | tee -a gpurun_out/${TAG}_pytest_vram_8gpu.log
tail -3 gpurun_out/${TAG}_pytest_vram_8gpu.log
run() { name=$1; shift; timeout 300 python tools/tier_sweep.py --gpus 8 "$@" > gpurun_out/${TAG}_tier_c5_1tib_${name}.json 2> gpurun_out/${TAG}_tier_c5_1tib_${name}.err; echo "== $name rc=$?"; tail -c 1300 gpurun_out/${TAG}_tier_c5_1tib_${name}.json; tail -2 gpurun_out/${TAG}_tier_c5_1tib_${name}.err; }
run fixed_ce_sender_a4 --fixed-frames --ahead 4 --laps 3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_8gpu.json 2> gpurun_out/${TAG}_bench_8gpu.err
echo "bench8 rc=$?"; tail -c 7000 gpurun_out/${TAG}_bench_8gpu.json; tail -5 gpurun_out/${TAG}_bench_8gpu.err
