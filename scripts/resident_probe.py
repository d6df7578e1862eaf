"""Launch list of one device-resident body tick (go1mpc_body_mpc_step_batch_resident_host_async) at B robots:
run under  ncu --metrics gpu__time_duration.sum --clock-control none --csv  to see the expand / pack copy kernels
beside the tick's own kernels.  Usage: python scripts/resident_probe.py [B]"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import quadrupedal_loco_b200 as q
from quadrupedal_loco_b200 import synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nh = 10
mpc = q.Go1Mpc(0)
d = synth.body_mpc_inputs(B, nh, seed=5)
rec = q.pack_body_inputs(nh, d["tick"], d["tx"], d["theta"], d["bstate"], d["x_warm"], d["refs"])
tx, xw, tick = q.split_body_record(nh, rec)
tx_d = torch.from_numpy(tx).cuda()
out_d = torch.zeros(B, q.body_out_stride(nh), dtype=torch.float64, device="cuda")
out_d[:, 18:18 + 2 * nh] = torch.from_numpy(xw).cuda()
ti = torch.from_numpy(tick).pin_memory()
to = torch.zeros(B, q.BODY_TICK_OUT, dtype=torch.float64).pin_memory()
dg = torch.zeros(B, q.body_diag_stride(nh), dtype=torch.int32).pin_memory()
torch.cuda.synchronize()
for _ in range(4):
    mpc.body_mpc_step_resident_host_async(nh, B, tx_d, out_d, ti.numpy(), to.numpy(), dg.numpy())
    mpc.synchronize()
print("status ok:", bool((dg.numpy()[:, 0] == 0).all()), "launches", mpc.launch_count)
