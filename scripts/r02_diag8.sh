#!/bin/bash
O=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --no-cpu-baseline --latency-samples 100 > $O/r02_diag8_peer.json 2> $O/r02_diag8.err

python - <<'PY'
import json
for n in ('peer',):
    for l in open(f'gpurun_out/r02_diag8_{n}.json'):
        if l.startswith('{'):
            d=json.loads(l); e=d['e2e']; print(n, 'value %.1fM'%(d['value']/1e6), 'e2e %.1fM'%(e['value']/1e6), 'e2e ms/step', round(e['ms_per_step'],4), 'wall', round(e['wall_ms_per_step'],4))
PY
