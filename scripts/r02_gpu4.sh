#!/bin/bash
O=gpurun_out
SWEEP_NH=20,40 python scripts/horizon_sweep.py > $O/r02_sweep2.log 2>&1
python -m pytest tests/test_gpu_body_modes.py tests/test_gpu_body.py -x -q -k "duo or dense or parity" >> $O/r02_sweep2.log 2>&1
cat $O/r02_sweep2.log | tail -12
