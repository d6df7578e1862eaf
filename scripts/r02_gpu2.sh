#!/bin/bash
# round-2 GPU pass 2: horizon sweep (cfg4) on the new any-horizon kernel, A/B against the dense kernel and the nh = 10 paths
O=gpurun_out
echo "== default (duo for nh 20/40, tri for 4/10)" > $O/r02_sweep.log
python scripts/horizon_sweep.py >> $O/r02_sweep.log 2>&1
echo "== GO1MPC_FORCE_GENERIC=1 (dense kernel)" >> $O/r02_sweep.log
GO1MPC_FORCE_GENERIC=1 python scripts/horizon_sweep.py >> $O/r02_sweep.log 2>&1
echo "== GO1MPC_BODY_MODE=duo (also nh 4/10)" >> $O/r02_sweep.log
GO1MPC_BODY_MODE=duo python scripts/horizon_sweep.py >> $O/r02_sweep.log 2>&1
cat $O/r02_sweep.log
