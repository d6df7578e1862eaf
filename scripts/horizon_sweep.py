"""cfg4: horizon sweep 10/20/40 of the body-inclination MPC tick (device-resident, one launch, CUDA events)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quadrupedal_loco_b200 as q
from quadrupedal_loco_b200 import synth
mpc = q.Go1Mpc(0); dev = torch.device("cuda", 0)
stream = torch.cuda.ExternalStream(mpc.stream, device=dev)
NHS = [int(v) for v in os.environ.get('SWEEP_NH', '4,10,20,40').split(',')]
BS = [int(v) for v in os.environ.get('SWEEP_B', '4096,32768').split(',')]
SCALE = float(os.environ.get('SWEEP_SCALE', '1.0'))
for nh in NHS:
    for B in BS:
        d = synth.body_mpc_inputs(B, nh, seed=nh, scale=SCALE)
        r = torch.from_numpy(q.pack_body_inputs(nh, d["tick"], d["tx"], d["theta"], d["bstate"], d["x_warm"], d["refs"])).to(dev)
        o = torch.zeros(B, q.body_out_stride(nh), dtype=torch.float64, device=dev)
        dg = torch.zeros(B, q.body_diag_stride(nh), dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        for _ in range(2): mpc.body_mpc_step(nh, B, r, o, dg)
        mpc.synchronize()
        with torch.cuda.stream(stream):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(5): mpc.body_mpc_step(nh, B, r, o, dg)
            e1.record(stream)
        e1.synchronize()
        ms = e0.elapsed_time(e1) / 5
        D = dg.cpu().numpy()
        print(f"nh={nh:2d} B={B:6d}: {ms * 1e3:9.1f} us/launch  {B / ms * 1e3:.3e} solves/s  mean adds {D[:, 3].mean():.2f}  "
              f"alg GFLOP/s {D[:, 9].astype(float).sum() / ms / 1e6:.0f}  status0 {np.mean(D[:, 0] == 0):.3f}")
