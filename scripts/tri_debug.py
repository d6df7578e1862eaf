import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import quadrupedal_loco_b200 as q
from quadrupedal_loco_b200 import synth
nh, B = 10, 256
mpc = q.Go1Mpc(0); dev = torch.device("cuda", 0)
d = synth.body_mpc_inputs(B, nh, seed=synth.SEED_CFG2)
rec = q.pack_body_inputs(nh, d["tick"], d["tx"], d["theta"], d["bstate"], d["x_warm"], d["refs"])
r = torch.from_numpy(rec).to(dev); o = torch.zeros(B, q.body_out_stride(nh), dtype=torch.float64, device=dev)
dg = torch.zeros(B, q.body_diag_stride(nh), dtype=torch.int32, device=dev)
torch.cuda.synchronize()
print("launching stages", os.environ.get("GO1MPC_TRI_STAGES"), flush=True)
t = time.time()
mpc.body_mpc_step(nh, B, r, o, dg); mpc.synchronize(); torch.cuda.synchronize()
print("done in", time.time() - t, "status0", int((dg[:, 0] == 0).sum()), "guard", mpc.body_guard_trips(), "handover", mpc.body_handover_total(), flush=True)
