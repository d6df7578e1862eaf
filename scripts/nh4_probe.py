import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import quadrupedal_loco_b200 as q
from quadrupedal_loco_b200 import synth
nh = 4
for mode in ("fast", "tri"):
    os.environ["GO1MPC_BODY_MODE"] = mode
    mpc = q.Go1Mpc(0); dev = torch.device("cuda", 0)
    stream = torch.cuda.ExternalStream(mpc.stream, device=dev)
    for B in (4096, 65536):
        d = synth.body_mpc_inputs(B, nh, seed=5)
        rec = torch.from_numpy(q.pack_body_inputs(nh, d["tick"], d["tx"], d["theta"], d["bstate"], d["x_warm"], d["refs"])).to(dev)
        o = torch.zeros(B, q.body_out_stride(nh), dtype=torch.float64, device=dev)
        for _ in range(5): mpc.body_mpc_step(nh, B, rec, o, None)
        mpc.synchronize()
        with torch.cuda.stream(stream):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(20): mpc.body_mpc_step(nh, B, rec, o, None)
            e1.record(stream)
        e1.synchronize()
        print(mode, "nh=4 B=%d: %.1f us/call" % (B, e0.elapsed_time(e1) / 20 * 1e3), flush=True)
    mpc.close()
