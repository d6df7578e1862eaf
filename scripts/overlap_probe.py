import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quadrupedal_loco_b200 as q
from quadrupedal_loco_b200 import synth
mpc = q.Go1Mpc(0); dev = torch.device("cuda", 0); nh, B, NR = 10, 4096, 16
d = synth.body_mpc_inputs(B * NR, nh, seed=1)
r = torch.from_numpy(q.pack_body_inputs(nh, d["tick"], d["tx"], d["theta"], d["bstate"], d["x_warm"], d["refs"])).to(dev).view(NR, B, -1)
o = torch.zeros(NR, B, q.body_out_stride(nh), dtype=torch.float64, device=dev)
pr = [r[k].data_ptr() for k in range(NR)]; po = [o[k].data_ptr() for k in range(NR)]
for ns in (1, 2, 4, 6, 8, 12):
    streams = [torch.cuda.Stream(device=dev) for _ in range(ns)]
    for k in range(ns):   # first call on a stream allocates its workspace
        mpc.lib.go1mpc_body_mpc_step_batch(mpc.h, nh, B, pr[0], po[0], None, streams[k].cuda_stream)
    torch.cuda.synchronize(); t = time.perf_counter()
    th = time.perf_counter()
    for i in range(400):
        mpc.lib.go1mpc_body_mpc_step_batch(mpc.h, nh, B, pr[i % NR], po[i % NR], None, streams[i % ns].cuda_stream)
    host = time.perf_counter() - th
    torch.cuda.synchronize(); el = time.perf_counter() - t
    print(f"body only, {ns} stream(s): {el / 400 * 1e6:.1f} us per 4096-batch -> {B * 400 / el / 1e6:.1f} M solves/s (host enqueue {host / 400 * 1e6:.1f} us per call)")
