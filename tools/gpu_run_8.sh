#!/bin/bash
# round-2 multi-GPU call (gpurun --gpus 8, charged 8x: keep it short): peer-tier tests, NVLink counters of one evict and one
# prefetch launch, the 1 TiB policy sweep in its variants, the 8-GPU bench line
mkdir -p gpurun_out
TAG=r02
nvidia-smi topo -m > gpurun_out/${TAG}_topo_8gpu.txt 2>&1
timeout 400 python -m pytest tests/test_gpu_vram.py -q --timeout 200 > gpurun_out/${TAG}_pytest_vram_8gpu.log 2>&1; echo "vram rc=$?" | tee -a gpurun_out/${TAG}_pytest_vram_8gpu.log
tail -3 gpurun_out/${TAG}_pytest_vram_8gpu.log
ncu --query-metrics 2>/dev/null | grep -i -E "nvlrx|nvltx" | head -40 > gpurun_out/${TAG}_nvl_metric_names.txt
CUDA_VISIBLE_DEVICES=0,1 timeout 300 ncu --metrics nvlrx__bytes.sum,nvltx__bytes.sum,gpu__time_duration.sum,lts__t_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -k regex:tfw_mover -c 6 --csv --log-file gpurun_out/${TAG}_peer_ncu.csv python tools/peer_ncu_probe.py 2 > gpurun_out/${TAG}_peer_ncu_probe.json 2> gpurun_out/${TAG}_peer_ncu.err
echo "ncu rc=$?"; tail -2 gpurun_out/${TAG}_peer_ncu.err; cat gpurun_out/${TAG}_peer_ncu_probe.json
CUDA_VISIBLE_DEVICES=0,1 timeout 200 python tools/peer_ncu_probe.py 8 > gpurun_out/${TAG}_peer_probe_8gib.json 2>&1; cat gpurun_out/${TAG}_peer_probe_8gib.json
for AH in 2 0 4; do
  timeout 400 python tools/tier_sweep.py --gpus 8 --ahead $AH > gpurun_out/${TAG}_tier_c5_1tib_ahead${AH}.json 2> gpurun_out/${TAG}_tier_c5_ahead${AH}.err; echo "sweep ahead=$AH rc=$?"; tail -c 1200 gpurun_out/${TAG}_tier_c5_1tib_ahead${AH}.json; tail -2 gpurun_out/${TAG}_tier_c5_ahead${AH}.err
done
timeout 400 python tools/tier_sweep.py --gpus 8 --ahead 2 --engine kernel > gpurun_out/${TAG}_tier_c5_1tib_ce.json 2> gpurun_out/${TAG}_tier_c5_ce.err; tail -c 1200 gpurun_out/${TAG}_tier_c5_1tib_ce.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_8gpu.json 2> gpurun_out/${TAG}_bench_8gpu.err
echo "bench8 rc=$?"; tail -c 5000 gpurun_out/${TAG}_bench_8gpu.json; tail -5 gpurun_out/${TAG}_bench_8gpu.err
