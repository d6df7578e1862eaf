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

# planner-only: step-timing SQP ticks dealt over ns streams (state refreshed by a D2D copy as in bench.py)
NRP = 8
stick, sst, sinp = synth.step_timing_inputs(B * NRP, mpc.step_default_state(), seed=2)
soa = lambda x, f: np.ascontiguousarray(x.reshape(NRP, B, f).transpose(0, 2, 1))
st_d = torch.from_numpy(soa(sst, q.STEP_STATE)).to(dev); si_d = torch.from_numpy(soa(sinp, q.STEP_IN)).to(dev)
tk_d = torch.from_numpy(stick.reshape(NRP, B).copy()).to(dev)
st_o = torch.zeros(NRP, B * q.STEP_STATE, dtype=torch.float64, device=dev)
so_d = torch.zeros(NRP, q.STEP_OUT, B, dtype=torch.float64, device=dev)
for ns in (1, 2, 4, 6, 8, 12):
    streams = [torch.cuda.Stream(device=dev) for _ in range(ns)]
    torch.cuda.synchronize(); t = time.perf_counter()
    for i in range(400):
        r_ = i % NRP; sp = streams[i % ns].cuda_stream
        mpc.lib.go1mpc_copy_device_async(mpc.h, st_o[r_].data_ptr(), st_d[r_].data_ptr(), B * q.STEP_STATE * 8, sp)
        mpc.lib.go1mpc_step_timing_step_batch(mpc.h, 3, B, tk_d[r_].data_ptr(), st_o[r_].data_ptr(), st_o[r_].data_ptr(), si_d[r_].data_ptr(), so_d[r_].data_ptr(), None, sp)
    torch.cuda.synchronize(); el = time.perf_counter() - t
    print(f"planner only, {ns} stream(s): {el / 400 * 1e6:.1f} us per 4096-batch")
