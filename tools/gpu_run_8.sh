#!/bin/bash
# round-2 final multi-GPU call (gpurun --gpus 8, charged 8x: keep it short): peer-tier tests, the 1 TiB policy sweep in its
# engines, the 8-GPU bench line
mkdir -p gpurun_out
TAG=r02f
export TFW_VS_DEBUG=1
timeout 300 python -m pytest tests/test_gpu_vram.py -q --timeout 200 > gpurun_out/${TAG}_pytest_vram_8gpu.log 2>&1; echo "vram rc=$?" | tee -a gpurun_out/${TAG}_pytest_vram_8gpu.log
tail -3 gpurun_out/${TAG}_pytest_vram_8gpu.log
run() { name=$1; shift; timeout 400 python tools/tier_sweep.py --gpus 8 "$@" > gpurun_out/${TAG}_tier_c5_1tib_${name}.json 2> gpurun_out/${TAG}_tier_c5_1tib_${name}.err; echo "== $name rc=$?"; tail -c 1500 gpurun_out/${TAG}_tier_c5_1tib_${name}.json; grep tfw_vspace gpurun_out/${TAG}_tier_c5_1tib_${name}.err; }
run ce_sender_a2 --ahead 2
run ce_sender_a3 --ahead 3
run kernel_sender_a2 --ahead 2 --engine kernel
run kernel_pull_a2 --ahead 2 --engine kernel --receiver-driven
run ce_pull_a2 --ahead 2 --receiver-driven
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_8gpu.json 2> gpurun_out/${TAG}_bench_8gpu.err
echo "bench8 rc=$?"; tail -c 6000 gpurun_out/${TAG}_bench_8gpu.json; tail -5 gpurun_out/${TAG}_bench_8gpu.err
