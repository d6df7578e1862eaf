"""Latency view of body_fast: B identical copies of the batch's slowest instance (most active-set
passes), one warp per SM, so the per-line stall samples of an ncu capture show the critical path."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quadrupedal_loco_b200 as q
from quadrupedal_loco_b200 import synth
nh, B = 10, int(sys.argv[1]) if len(sys.argv) > 1 else 148
mpc = q.Go1Mpc(0); dev = torch.device("cuda", 0)
d = synth.body_mpc_inputs(4096, nh, seed=1)
rec = q.pack_body_inputs(nh, d["tick"], d["tx"], d["theta"], d["bstate"], d["x_warm"], d["refs"])
out = np.zeros((4096, q.body_out_stride(nh))); diag = np.zeros((4096, q.body_diag_stride(nh)), np.int32)
mpc.body_mpc_step_host(nh, 4096, rec, out, diag)
worst = int(np.argmax(diag[:, 8]))
print("slowest instance", worst, "passes", diag[worst, 8], "iters", diag[worst, 2:6])
r = torch.from_numpy(np.tile(rec[worst], (B, 1))).to(dev); o = torch.zeros(B, q.body_out_stride(nh), dtype=torch.float64, device=dev)
stream = torch.cuda.ExternalStream(mpc.stream, device=dev)
torch.cuda.synchronize()
for _ in range(3): mpc.body_mpc_step(nh, B, r, o, None)
mpc.synchronize()
with torch.cuda.stream(stream):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(10): mpc.body_mpc_step(nh, B, r, o, None)
    e1.record(stream)
e1.synchronize()
us = e0.elapsed_time(e1) / 10 * 1e3
print(f"B={B}: {us:.1f} us per launch -> {us * 1.965e3 / diag[worst, 8]:.0f} cycles per pass (setup included)")
