#!/bin/bash
for cfg in "65536 2" "65536 3" "65536 4" "8192 8" "8192 12" "16384 4" "16384 8"; do set -- $cfg
python bench.py --batch $1 --lanes $2 --no-cpu-baseline --no-e2e --latency-samples 20 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('B $1 lanes $2 value %.1fM ms/step %.4f'%(d['value']/1e6,d['ms_per_step']))
"
done
