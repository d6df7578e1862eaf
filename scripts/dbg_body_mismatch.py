"""Debug aid: list body-MPC instances where the GPU and the oracle disagree."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quadrupedal_loco_b200 as q
from quadrupedal_loco_b200 import synth
from tests import oracle_lib
from tests.test_gpu_body import run_gpu, run_oracle

nh = int(sys.argv[1]) if len(sys.argv) > 1 else 10
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
seed = int(sys.argv[3], 0) if len(sys.argv) > 3 else synth.SEED_CFG3
scale = float(sys.argv[4]) if len(sys.argv) > 4 else 2.0
mpc = q.Go1Mpc(0); orc = oracle_lib.Oracle()
d = synth.body_mpc_inputs(B, nh, seed=seed, scale=scale)
out, diag = run_gpu(mpc, nh, d)
r = run_oracle(orc, nh, d)
A0 = q.BODY_DIAG_ACTIVE
bad = np.nonzero((diag[:, 0] != r["status"]) | (diag[:, 1] != r["nactive"]) | (diag[:, 2:6] != r["iters"]).any(axis=1))[0]
print("mismatching instances:", len(bad), "of", B)
for b in bad[:12]:
    k = max(diag[b, 1], r["nactive"][b])
    print(f"b={b} gpu: st={diag[b,0]} na={diag[b,1]} it={diag[b,2:6]} l2a={diag[b,8]} A={diag[b,A0:A0+k]}")
    print(f"      orc: st={r['status'][b]} na={r['nactive'][b]} it={r['iters'][b]} A={r['active'][b,:k]}")
    xs = np.abs(out[b, 18:18 + 2 * nh] - r["x"][b]).max() / max(1, np.abs(r["x"][b]).max())
    print(f"      x rel err {xs:.3e}  theta={d['theta'][b]}  cost gpu={out[b,18+2*nh]:.6e}")
ok = np.setdiff1d(np.arange(B), bad)
scale_ = np.maximum(1.0, np.abs(r["x"]).max(axis=1, keepdims=True))
err = (np.abs(out[:, 18:18 + 2 * nh] - r["x"]) / scale_)
print("max rel err over matching instances:", err[ok].max(), " over all:", err.max())
