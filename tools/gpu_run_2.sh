#!/bin/bash
# round-2 iteration call (gpurun --gpus 2): VMM call costs, peer-tier tests, the policy sweep at 2 GPUs in its variants
mkdir -p gpurun_out
TAG=r02
timeout 120 tools/vmm_lab 2 > gpurun_out/${TAG}_vmm_lab_2gpu.jsonl 2>&1; cat gpurun_out/${TAG}_vmm_lab_2gpu.jsonl
timeout 400 python -m pytest tests/test_gpu_vram.py -q --timeout 200 > gpurun_out/${TAG}_pytest_vram_2gpu.log 2>&1; echo "vram rc=$?" | tee -a gpurun_out/${TAG}_pytest_vram_2gpu.log
grep -E "^E  |FAILED|passed|failed" gpurun_out/${TAG}_pytest_vram_2gpu.log | head -20
timeout 200 python tools/peer_ncu_probe.py 8 > gpurun_out/${TAG}_peer_probe_8gib_2gpu.json 2>&1; cat gpurun_out/${TAG}_peer_probe_8gib_2gpu.json
timeout 200 python tools/peer_ncu_probe.py 8 --ce > gpurun_out/${TAG}_peer_probe_8gib_2gpu_ce.json 2>&1; cat gpurun_out/${TAG}_peer_probe_8gib_2gpu_ce.json
for AH in 2 0 4; do
  timeout 300 python tools/tier_sweep.py --gpus 2 --ahead $AH > gpurun_out/${TAG}_tier_2gpu_ahead${AH}.json 2> gpurun_out/${TAG}_tier_2gpu_ahead${AH}.err; echo "sweep ahead=$AH rc=$?"; tail -c 900 gpurun_out/${TAG}_tier_2gpu_ahead${AH}.json; tail -2 gpurun_out/${TAG}_tier_2gpu_ahead${AH}.err
done
timeout 300 python tools/tier_sweep.py --gpus 2 --ahead 2 --copy-engine > gpurun_out/${TAG}_tier_2gpu_ce.json 2> gpurun_out/${TAG}_tier_2gpu_ce.err; tail -c 900 gpurun_out/${TAG}_tier_2gpu_ce.json; tail -2 gpurun_out/${TAG}_tier_2gpu_ce.err
