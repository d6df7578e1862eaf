#!/bin/bash
O=gpurun_out
for B in 65536 8192; do
for occ in 1 2 3; do
  for L in 2 4 8 16; do
    GO1MPC_TRI_OCC=$occ python bench.py --batch $B --lanes $L --no-cpu-baseline --no-e2e --latency-samples 20 > $O/tmp_occ.json 2>/dev/null
    python - <<PY
import json
for l in open('$O/tmp_occ.json'):
    if l.startswith('{'):
        d=json.loads(l); print('B $B occ $occ lanes $L value %.1fM ms/step %.4f'%(d['value']/1e6, d['ms_per_step']))
PY
  done
done
done
