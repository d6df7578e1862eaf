#!/bin/bash
# round-2 GPU pass 1: full GPU test suite, bench (cfg3 / cfg2, lane sweep), launch list, ncu full captures of the top kernels
set -x
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02_pytest1.log 2>&1; echo "pytest rc=$?"
python bench.py > $O/r02_c1.json 2> $O/r02_c1.err; echo "bench rc=$?"
python bench.py --config cfg2 --no-cpu-baseline > $O/r02_c1_cfg2.json 2> $O/r02_c1_cfg2.err
for L in 1 4 8; do python bench.py --lanes $L --no-cpu-baseline --no-e2e --latency-samples 50 > $O/r02_c1_L$L.json 2>/dev/null; done
for B in 32768 8192; do for L in 4 8 16; do python bench.py --batch $B --lanes $L --no-cpu-baseline --no-e2e --latency-samples 50 > $O/r02_c1_B${B}_L$L.json 2>/dev/null; done; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --latency-samples 10 > $O/r02_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'tri_solve|sqp|height|post|tri_setup|tri_merge' -c 24 -o $O/r02_full1 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-graph --latency-samples 4 > $O/r02_ncu_full.log 2>&1
ls -la $O | tail -5
