#!/bin/bash
O=gpurun_out
SWEEP_NH=20 SWEEP_B=8192 ncu --set full --clock-control none --import-source on -k regex:body_duo -c 2 -o $O/r02_duo_nh20 python scripts/horizon_sweep.py > $O/r02_duo_ncu.log 2>&1
tail -3 $O/r02_duo_ncu.log
