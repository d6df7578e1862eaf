#!/bin/bash
O=gpurun_out
run() { python bench.py --no-cpu-baseline --no-e2e --latency-samples 100 > $O/tmp_v.json 2>/dev/null; python - <<PY
import json
for l in open('$O/tmp_v.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$1 value %.1fM ms/step %.4f body_alone %.4f'%(d['value']/1e6, d['ms_per_step'], d['kernels_ms']['body_tick_alone']))
PY
}
run default
for v in "-DGO1_TRI_WPC=2 -DGO1_TRI_MAXNREG=144" "-DGO1_TRI_WPC=2 -DGO1_TRI_MAXNREG=136" "-DGO1_TRI_WPC=2 -DGO1_TRI_WARPS=12" "-DGO1_TRI_WPC=1 -DGO1_TRI_MAXNREG=144"; do
  GO1MPC_NVCC_EXTRA="$v" python -c "from quadrupedal_loco_b200 import _build; _build.build(force=True)" > /dev/null 2>&1
  run "$v"
  python -m pytest tests/test_gpu_body_modes.py -x -q -k "tri and ragged" 2>&1 | tail -1
done
