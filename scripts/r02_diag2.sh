#!/bin/bash
O=gpurun_out
for G in peer nccl; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --no-cpu-baseline --latency-samples 100 --gather $G > $O/r02_diag2_$G.json 2> $O/r02_diag2_$G.err
tail -c 300 $O/r02_diag2_$G.err
done
python - <<'PY'
import json
for n in ('peer','nccl'):
    for l in open(f'gpurun_out/r02_diag2_{n}.json'):
        if l.startswith('{'):
            d=json.loads(l); e=d['e2e']; print(n, 'value %.1fM'%(d['value']/1e6), 'e2e %.1fM'%(e['value']/1e6), 'gather', e.get('gather'), 'gather_ms', e.get('gather_ms'), 'lat', e['latency_ms']['p99'])
PY
