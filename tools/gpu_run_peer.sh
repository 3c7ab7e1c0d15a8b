#!/bin/bash
# round-2 multi-GPU call (gpurun --gpus 2): peer-tier tests, NVLink counters of one evict and one prefetch launch, the 2-GPU bench
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_2gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_vram.py -q --timeout 300 > gpurun_out/r02_pytest_vram_2gpu.log 2>&1; echo "vram rc=$?" | tee -a gpurun_out/r02_pytest_vram_2gpu.log
tail -3 gpurun_out/r02_pytest_vram_2gpu.log
ncu --query-metrics 2>/dev/null | grep -i -E "^nvl|nvlrx|nvltx" | head -40 > gpurun_out/r02_nvl_metric_names.txt
timeout 600 ncu --metrics nvlrx__bytes.sum,nvltx__bytes.sum,gpu__time_duration.sum,lts__t_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -k regex:tfw_mover -c 6 --csv --log-file gpurun_out/r02_peer_ncu.csv python tools/peer_ncu_probe.py 2 > gpurun_out/r02_peer_ncu_probe.json 2> gpurun_out/r02_peer_ncu.err
echo "ncu rc=$?"; tail -3 gpurun_out/r02_peer_ncu.err; cat gpurun_out/r02_peer_ncu_probe.json
timeout 300 python tools/peer_ncu_probe.py 8 > gpurun_out/r02_peer_probe_8gib.json 2>&1; cat gpurun_out/r02_peer_probe_8gib.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err
echo "bench2 rc=$?"; tail -c 3000 gpurun_out/r02_bench_2gpu.json; tail -5 gpurun_out/r02_bench_2gpu.err
