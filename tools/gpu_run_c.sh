#!/bin/bash
# round-2 GPU call C (1 GPU): fixed-frames tiering (host tier), where unpark spends its time, what stalls a C3 tenant
mkdir -p gpurun_out
TAG=r02c
timeout 900 python -m pytest tests/test_gpu_vram.py tests/test_worker_binary.py tests/test_gpu_isolation.py -m gpu -q --timeout 300 > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log; tail -12 gpurun_out/${TAG}_pytest.log
timeout 120 python -m pytest tests/test_gpu_isolation.py -q -k parked -s 2>&1 | grep -E "park|tfw\]" > gpurun_out/${TAG}_park.txt; cat gpurun_out/${TAG}_park.txt
for P in 1 0; do
  TFW_BRIDGE_PACED=$P timeout 200 python tools/limiter_c3.py --seconds 12 --workers 4 --limit 25 --feedback device > gpurun_out/${TAG}_c3_paced$P.json 2> gpurun_out/${TAG}_c3_paced$P.err
  echo "c3 paced=$P rc=$?"; tail -c 3000 gpurun_out/${TAG}_c3_paced$P.json
done
for V in "" "--fixed-frames" "--fixed-frames --ahead 3"; do
  N=$(echo "lru$V" | tr -d ' -')
  timeout 300 python tools/tier_sweep.py --gpus 1 --laps 3 $V > gpurun_out/${TAG}_tier_c4_$N.json 2> gpurun_out/${TAG}_tier_c4_$N.err
  echo "c4 $N rc=$?"; tail -c 1500 gpurun_out/${TAG}_tier_c4_$N.json; tail -3 gpurun_out/${TAG}_tier_c4_$N.err
done
ls -la gpurun_out/ | grep $TAG
