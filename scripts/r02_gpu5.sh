#!/bin/bash
O=gpurun_out
for occ in 1 2 3; do
  for B in 8192 65536; do
    GO1MPC_TRI_OCC=$occ python bench.py --batch $B --no-cpu-baseline --no-e2e --latency-samples 50 > $O/r02_occ${occ}_B$B.json 2>/dev/null
    python - <<PY
import json
for l in open('$O/r02_occ${occ}_B$B.json'):
    if l.startswith('{'):
        d=json.loads(l); print('occ $occ B $B value %.1fM ms/step %.4f body_alone %.4f'%(d['value']/1e6, d['ms_per_step'], d['kernels_ms']['body_tick_alone']))
PY
  done
done
