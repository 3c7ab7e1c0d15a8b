#!/bin/bash
# round-2 GPU call E (2 GPUs): fixed-frames tiering over NVLink -- parity tests, then the policy sweep next to the LRU engine
mkdir -p gpurun_out
TAG=r02e
timeout 300 python -m pytest tests/test_gpu_vram.py -m gpu -q --timeout 200 > gpurun_out/${TAG}_pytest_vram_2gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_vram_2gpu.log; tail -6 gpurun_out/${TAG}_pytest_vram_2gpu.log
run() { N=$1; shift; timeout 200 python tools/tier_sweep.py --gpus 2 --laps 4 "$@" > gpurun_out/${TAG}_tier_2gpu_$N.json 2> gpurun_out/${TAG}_tier_2gpu_$N.err
  echo "== $N rc=$?"; python - gpurun_out/${TAG}_tier_2gpu_$N.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d[k] for k in ('engine','copies_driven_by','va_repointed','prefetch_ahead','lap_seconds','prefetch_frac_of_nvlink_nominal_900','median_lap_frac_of_nvlink_nominal_900','host_stall_ms_per_lap','vmm_ms_per_lap')})
except Exception as e: print('failed', e)
PY
  tail -2 gpurun_out/${TAG}_tier_2gpu_$N.err; }
run fixed_ce_sender_a2 --fixed-frames
run fixed_ce_sender_a4 --fixed-frames --ahead 4
run fixed_kernel_sender_a2 --fixed-frames --engine kernel
run fixed_ce_pull_a2 --fixed-frames --receiver-driven
run lru_ce_sender_a2
