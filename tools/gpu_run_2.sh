#!/bin/bash
# round-2 iteration call (gpurun --gpus 2): VMM call costs, the policy sweep at 2 GPUs: who drives the copies, and how far ahead
mkdir -p gpurun_out
TAG=r02
timeout 120 tools/vmm_lab 2 > gpurun_out/${TAG}_vmm_lab_2gpu.jsonl 2>&1; head -4 gpurun_out/${TAG}_vmm_lab_2gpu.jsonl
timeout 300 python -m pytest tests/test_gpu_vram.py -q --timeout 200 2>&1 | tail -3
run() { name=$1; shift; timeout 300 python tools/tier_sweep.py --gpus 2 "$@" > gpurun_out/${TAG}_tier_2gpu_${name}.json 2> gpurun_out/${TAG}_tier_2gpu_${name}.err; echo "== $name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_tier_2gpu_${name}.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("lap_seconds","prefetch_GBps_into_home_gpu","prefetch_frac_of_nvlink_nominal_900","hits_inflight","host_stall_ms_per_lap")})
except Exception as e:
    print("no result", e); print(open("gpurun_out/${TAG}_tier_2gpu_${name}.err").read()[-400:])
PY
}
run pull_a2 --ahead 2 --laps 3
run pull_a1 --ahead 1 --laps 3
run pull_a3 --ahead 3 --laps 3
run sender_a2 --ahead 2 --laps 3 --sender-driven
run sender_a1 --ahead 1 --laps 3 --sender-driven
run ce_pull_a2 --ahead 2 --laps 3 --copy-engine
run ce_sender_a2 --ahead 2 --laps 3 --copy-engine --sender-driven
run homepush_a2 --ahead 2 --laps 3 --push-evict
for v in "" "--sender" "--ce" "--ce --sender"; do timeout 120 python tools/peer_ncu_probe.py 8 $v 2>&1 | tail -1; done > gpurun_out/${TAG}_peer_probe_variants_2gpu.jsonl; cat gpurun_out/${TAG}_peer_probe_variants_2gpu.jsonl
