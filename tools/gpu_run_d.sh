#!/bin/bash
# round-2 GPU call D (1 GPU): fixed-frames sweep with a bounded pipeline, continuous park/unpark ring, C3 with averaged utilisation samples
mkdir -p gpurun_out
TAG=r02d
timeout 600 python -m pytest tests/test_gpu_vram.py tests/test_gpu_isolation.py -m gpu -q --timeout 300 > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log; tail -12 gpurun_out/${TAG}_pytest.log
timeout 120 python -m pytest tests/test_gpu_isolation.py -q -k parked -s 2>&1 | grep -E "park|tfw\]" > gpurun_out/${TAG}_park.txt; cat gpurun_out/${TAG}_park.txt
for S in 0 1; do
  E=""; [ $S = 1 ] && E="TF_UTIL_SINGLE_SAMPLE=1"
  env $E timeout 200 python tools/limiter_c3.py --seconds 12 --workers 4 --limit 25 --feedback device > gpurun_out/${TAG}_c3_single$S.json 2> gpurun_out/${TAG}_c3_single$S.err
  echo "c3 single_sample=$S rc=$?"; python - gpurun_out/${TAG}_c3_single$S.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k:v for k,v in d.items() if k not in('workers','config')})
PY
done
for V in "--fixed-frames" "--fixed-frames --ahead 4"; do
  N=$(echo "$V" | tr -d ' -')
  timeout 300 python tools/tier_sweep.py --gpus 1 --laps 3 $V > gpurun_out/${TAG}_tier_c4_$N.json 2> gpurun_out/${TAG}_tier_c4_$N.err
  echo "c4 $N rc=$?"; tail -c 1500 gpurun_out/${TAG}_tier_c4_$N.json; tail -3 gpurun_out/${TAG}_tier_c4_$N.err
done
