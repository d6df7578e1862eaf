#!/bin/bash
O=gpurun_out
ncu --metrics gpu__time_duration.sum,launch__registers_per_thread,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none --launch-skip 11000 -c 60 --csv --log-file $O/r02_chain_launches.csv python scripts/chain_probe.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/r02_chain_launches.csv')))
hi=next(i for i,r in enumerate(rows) if r and r[0]=='ID')
h=rows[hi]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); gi=h.index('Grid Size')
agg=collections.OrderedDict()
for r in rows[hi+1:]:
    if len(r)<=vi: continue
    key=(r[ki].split('(')[0], r[gi]); a=agg.setdefault(key, collections.defaultdict(list)); a[r[mi]].append(float(r[vi].replace(',','')))
for (n,g),a in agg.items():
    print(f"{n:40s} grid {g:14s} n={len(a['gpu__time_duration.sum']):3d} " + ' '.join(f"{k.split('.')[0].split('__')[-1]}={sum(v)/len(v):.4g}" for k,v in a.items()))
PY
