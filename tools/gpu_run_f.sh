#!/bin/bash
# round-2 GPU call F (1 GPU): C3 after the pacing fix -- both feedback signals, two burst lengths
mkdir -p gpurun_out
TAG=r02f1
for FB in device process; do for Q in 50 20; do
  TFW_BRIDGE_QUANTUM_MS=$Q timeout 200 python tools/limiter_c3.py --seconds 10 --workers 4 --limit 25 --feedback $FB > gpurun_out/${TAG}_c3_${FB}_q$Q.json 2> gpurun_out/${TAG}_c3_${FB}_q$Q.err
  echo "c3 $FB quantum=$Q rc=$?"; python - gpurun_out/${TAG}_c3_${FB}_q$Q.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k:v for k,v in d.items() if k not in('workers','config','ticks_t_util_nsamples')})
print('  util', [t[1] for t in d['ticks_t_util_nsamples']])
print('  w0 stalls', [(s['at_s'],s['batch_ms'],s['rate']) for s in d['workers'][0]['stalls']])
PY
done; done
timeout 300 python tools/tier_sweep.py --gpus 1 --laps 2 --fixed-frames --ahead 4 > gpurun_out/${TAG}_tier_c4_fixed_a4.json 2> gpurun_out/${TAG}_tier_c4_fixed_a4.err
echo "c4 fixed a4 rc=$?"; tail -c 900 gpurun_out/${TAG}_tier_c4_fixed_a4.json; tail -2 gpurun_out/${TAG}_tier_c4_fixed_a4.err
