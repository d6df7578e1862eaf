"""Where does the e2e step time go?  enqueue-only host time vs completion time of the pipelined host entries."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quadrupedal_loco_b200 as q
from quadrupedal_loco_b200 import synth
nh, B, NR, K = 10, 4096, 48, 48
mpc = q.Go1Mpc(0); lib, hh = mpc.lib, mpc.h
dev = torch.device("cuda", 0)
in_s, out_s, dg_s = q.body_in_stride(nh), q.body_out_stride(nh), q.body_diag_stride(nh)
big = synth.body_mpc_inputs(B * NR, nh, seed=1)
rec = torch.from_numpy(q.pack_body_inputs(nh, big["tick"], big["tx"], big["theta"], big["bstate"], big["x_warm"], big["refs"])).pin_memory().view(NR, B, in_s).numpy()
out = torch.zeros(NR, B, out_s, dtype=torch.float64).pin_memory().numpy(); dg = torch.zeros(NR, B, dg_s, dtype=torch.int32).pin_memory().numpy()
tk, st, si = synth.step_timing_inputs(B * NR, mpc.step_default_state(), seed=1)
soa = lambda x, f: np.ascontiguousarray(x.reshape(NR, B, f).transpose(0, 2, 1))
st_d = torch.from_numpy(soa(st, q.STEP_STATE)).to(dev); sto_d = torch.zeros_like(st_d)
si_h = torch.from_numpy(soa(si, q.STEP_IN)).pin_memory().numpy(); tk_h = torch.from_numpy(tk.reshape(NR, B).copy()).pin_memory().numpy()
so_h = torch.zeros(NR, q.STEP_OUT, B, dtype=torch.float64).pin_memory().numpy(); sd_h = torch.zeros(NR, q.STEP_DIAG, B, dtype=torch.int32).pin_memory().numpy()
P = lambda a: [a[r].ctypes.data for r in range(NR)]
Hin, Hout, Hdg, Htk, Hsi, Hso, Hsd = P(rec), P(out), P(dg), P(tk_h), P(si_h), P(so_h), P(sd_h)
Dst = [st_d[r].data_ptr() for r in range(NR)]; Dsto = [sto_d[r].data_ptr() for r in range(NR)]
def run(body=True, sqp=True, diag=True):
    mpc.synchronize(); t0 = time.perf_counter()
    for i in range(K):
        r = i % NR
        if sqp: lib.go1mpc_step_timing_step_batch_host_async(hh, 3, B, Htk[r], Dst[r], Dsto[r], Hsi[r], Hso[r], Hsd[r] if diag else None)
        if body: lib.go1mpc_body_mpc_step_batch_host_async(hh, nh, B, Hin[r], Hout[r], Hdg[r] if diag else None)
    t1 = time.perf_counter(); mpc.synchronize(); t2 = time.perf_counter()
    return (t1 - t0) / K * 1e6, (t2 - t0) / K * 1e6
for _ in range(2): run()
print("both      enqueue %.1f us/step, total %.1f us/step" % run())
print("body only enqueue %.1f us/step, total %.1f us/step" % run(sqp=False))
print("sqp only  enqueue %.1f us/step, total %.1f us/step" % run(body=False))
print("both, no diag: enqueue %.1f, total %.1f" % run(diag=False))
