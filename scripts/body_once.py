"""One body-MPC launch at batch B (argv[1], default 65536) after two warm-up launches: the target of an ncu capture."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quadrupedal_loco_b200 as q
from quadrupedal_loco_b200 import synth
nh, B = 10, int(sys.argv[1]) if len(sys.argv) > 1 else 65536
mpc = q.Go1Mpc(0); dev = torch.device("cuda", 0)
d = synth.body_mpc_inputs(B, nh, seed=synth.SEED_CFG2)
rec = q.pack_body_inputs(nh, d["tick"], d["tx"], d["theta"], d["bstate"], d["x_warm"], d["refs"])
r = torch.from_numpy(rec).to(dev); o = torch.zeros(B, q.body_out_stride(nh), dtype=torch.float64, device=dev)
dg = torch.zeros(B, q.body_diag_stride(nh), dtype=torch.int32, device=dev)
torch.cuda.synchronize()
for _ in range(3): mpc.body_mpc_step(nh, B, r, o, dg)
mpc.synchronize()
print("ok", int((dg[:, 0] == 0).sum()))
