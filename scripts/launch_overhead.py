import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quadrupedal_loco_b200 as q
from quadrupedal_loco_b200 import synth
mpc = q.Go1Mpc(0); dev = torch.device("cuda", 0); nh = 10
for B in (1, 4096):
    d = synth.body_mpc_inputs(B, nh, seed=1)
    r = torch.from_numpy(q.pack_body_inputs(nh, d["tick"], d["tx"], d["theta"], d["bstate"], d["x_warm"], d["refs"])).to(dev)
    o = torch.zeros(B, q.body_out_stride(nh), dtype=torch.float64, device=dev)
    pr, po = r.data_ptr(), o.data_ptr()
    torch.cuda.synchronize()
    for _ in range(10): mpc.lib.go1mpc_body_mpc_step_batch(mpc.h, nh, B, pr, po, None, None)
    mpc.synchronize()
    t = time.perf_counter()
    for _ in range(300): mpc.lib.go1mpc_body_mpc_step_batch(mpc.h, nh, B, pr, po, None, None)
    t1 = time.perf_counter(); mpc.synchronize(); t2 = time.perf_counter()
    print(f"B={B}: enqueue {(t1 - t) / 300 * 1e6:.1f} us/call, total {(t2 - t) / 300 * 1e6:.1f} us/call")
t = time.perf_counter()
for _ in range(1000): mpc.lib.go1mpc_device(mpc.h)
print("trivial ctypes call", (time.perf_counter() - t) / 1000 * 1e6, "us")
