#!/bin/bash
O=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'step_sqp|step_height|step_finish' --launch-skip 30 -c 3 -o $O/r02_sqp_b python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-graph --latency-samples 4 > $O/r02_ncu_sqp.log 2>&1
tail -2 $O/r02_ncu_sqp.log
