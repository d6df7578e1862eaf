import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quadrupedal_loco_b200 as q
from quadrupedal_loco_b200 import synth
from tests import oracle_lib
from tests.test_gpu_step import gpu_tick
mpc = q.Go1Mpc(0); orc = oracle_lib.Oracle(); cfg = orc.step_cfg(3)
np.set_printoptions(precision=12, linewidth=200)
for B, seed in ((1, 1), (33, 33)):
    tick, st, inp = synth.step_timing_inputs(B, mpc.step_default_state(), seed=seed, amp=0.5)
    for K in (1, 2, 3):
        cfgk = orc.step_cfg(K)
        go, gs, gd = gpu_tick(mpc, tick, st, inp, K)
        os_ = st.copy(); oo, od = orc.step_tick_batch(cfgk, tick, os_, inp)
        err = np.abs(gs - os_).max(axis=1)
        bad = np.nonzero(err > 1e-9)[0]
        print(f"B={B} K={K}: bad instances {bad[:10]} of {B}")
        for b in bad[:2]:
            print("  tick", tick[b], "diag gpu", gd[b, :16], "\n  diag orc", od[b, :16])
            print("  v gpu", gs[b, 195:199], "\n  v orc", os_[b, 195:199], "\n  v in ", st[b, 195:199])
            print("  ts gpu", gs[b, gd[b,0]-1], "ts orc", os_[b, od[b,0]-1])
