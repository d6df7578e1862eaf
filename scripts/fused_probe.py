"""cfg5: throughput of go1mpc_fused_tick_batch (planner tick -> swing foot -> body MPC -> servo IK), device-resident,
CUDA events, calls in flight on S streams."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quadrupedal_loco_b200 as q
from tests.test_gpu_fused import make_inputs
mpc = q.Go1Mpc(0); dev = torch.device("cuda", 0)
for B, S in ((4096, 8), (65536, 2), (131072, 2)):
    f64 = dict(dtype=torch.float64, device=dev); i32 = dict(dtype=torch.int32, device=dev)
    slots = []
    for s in range(S):
        I = make_inputs(mpc, B, seed=100 + s)
        slots.append(dict(I=I, st0=I["state"].clone(), out38=torch.zeros(q.STEP_OUT, B, **f64), out18=torch.zeros(18, B, **f64),
                          bout=torch.zeros(B, q.body_out_stride(10), **f64), theta=torch.zeros(3, B, **f64),
                          stream=torch.cuda.Stream(device=dev)))
    def call(k):
        z = slots[k % S]; I = z["I"]
        mpc.fused_tick(B, 3, I["tick"], I["state"], I["sin"], z["out38"], I["foot"], z["out18"], 10, I["rec"], z["bout"], 102, 0.7,
                       I["homing"], I["q"], z["theta"], stream=z["stream"].cuda_stream)
    for k in range(2 * S): call(k)
    torch.cuda.synchronize()
    n = 200 if B <= 4096 else 30
    t = time.perf_counter()
    for k in range(n): call(k)
    torch.cuda.synchronize(); el = time.perf_counter() - t
    print(f"fused tick B={B}, {S} streams: {el / n * 1e6:.1f} us per call -> {B * n / el / 1e6:.1f} M robot ticks/s ({4 * B * n / el / 1e6:.0f} M QP solves/s + 4 leg IKs per robot)")
