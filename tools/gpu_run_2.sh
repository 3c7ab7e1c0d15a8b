#!/bin/bash
# round-2 iteration call (gpurun --gpus 2): the policy sweep at 2 GPUs with lazily mapped peer regions: who drives the copies, with which engine
mkdir -p gpurun_out
TAG=r02
run() { name=$1; shift; timeout 300 python tools/tier_sweep.py --gpus 2 "$@" > gpurun_out/${TAG}_tier_2gpu_${name}.json 2> gpurun_out/${TAG}_tier_2gpu_${name}.err; echo "== $name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_tier_2gpu_${name}.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("lap_seconds","prefetch_GBps_into_home_gpu","prefetch_frac_of_nvlink_nominal_900","hits_inflight","host_stall_ms_per_lap","vmm_ms_per_lap","backings_created_in_sweeps","backings_destroyed_in_sweeps")})
except Exception as e:
    print("no result", e); print(open("gpurun_out/${TAG}_tier_2gpu_${name}.err").read()[-400:])
PY
}
export TFW_VS_DEBUG=1
show() { python - <<PY
import json
d=json.loads(open("gpurun_out/r02_tier_2gpu_$1.json").read().strip().splitlines()[-1]); print(d["laps"])
PY
grep tfw_vspace gpurun_out/r02_tier_2gpu_$1.err; }
run ce_pull_a2 --ahead 2 --laps 4 --copy-engine; show ce_pull_a2
TFW_VS_REMAP_LATE=1 run ce_pull_a2_late --ahead 2 --laps 4 --copy-engine; show ce_pull_a2_late
TFW_VS_REMAP_LATE=1 run ce_pull_a3_late --ahead 3 --laps 4 --copy-engine; show ce_pull_a3_late
TFW_VS_REMAP_LATE=1 TFW_VS_PEER_CTAS=2 run sender_a2_ctas2_late --ahead 2 --laps 4 --sender-driven; show sender_a2_ctas2_late
TFW_VS_REMAP_LATE=1 TFW_VS_PEER_CTAS=2 run pull_a2_ctas2_late --ahead 2 --laps 4; show pull_a2_ctas2_late
TFW_VS_REMAP_LATE=1 run ce_sender_a2_late --ahead 2 --laps 4 --copy-engine --sender-driven; show ce_sender_a2_late
timeout 300 python -m pytest tests/test_gpu_vram.py -q --timeout 200 2>&1 | tail -3
