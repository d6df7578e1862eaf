#!/bin/bash
O=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/r02_launches_b8192.csv python bench.py --batch 8192 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-graph --latency-samples 10 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/r02_launches_b8192.csv')))
hi=next(i for i,r in enumerate(rows) if r and r[0]=='ID')
h=rows[hi]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); gi=h.index('Grid Size')
agg=collections.OrderedDict()
for r in rows[hi+1:]:
    if len(r)<=vi: continue
    n=r[ki].split('(')[0]; v=float(r[vi].replace(',',''))
    a=agg.setdefault(n,[0,0.0,r[gi]]); a[0]+=1; a[1]+=v
for n,a in agg.items(): print(f"{n:50s} {a[0]:5d} launches {a[1]/a[0]/1000:9.1f} us avg grid {a[2]}")
PY
