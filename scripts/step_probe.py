"""Planner tick alone: mean duration of a lone call (CUDA events, out of place from a pristine state, inputs rotating
through 4 batches) for batch sizes argv[1:] (default 4096 65536), on bench.py's cfg2 and cfg3 planner mixes.
GO1MPC_STEP_MODE=thread1 selects round 1's single kernel, GO1MPC_LIB another build."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quadrupedal_loco_b200 as q
from quadrupedal_loco_b200 import synth
Bs = [int(x) for x in sys.argv[1:]] or [4096, 65536]
dev = torch.device("cuda", 0)
for name, amp, lam in (("cfg2", 1.0, (0, 0, 0, 0)), ("cfg3", 2.0, (0.25, 0.001, 0.025, 0.001))):
    mpc = q.Go1Mpc(0, {"step": {"lamda": lam}})
    for B in Bs:
        nrot = 4
        sl = []
        for r in range(nrot):
            tk, st, si = synth.step_timing_inputs(B, mpc.step_default_state(), seed=100 + r, amp=amp, push_x=0.4, push_y=0.75, p_hi=16)
            sl.append((torch.from_numpy(tk.copy()).to(dev), torch.from_numpy(np.ascontiguousarray(st.T)).to(dev), torch.from_numpy(np.ascontiguousarray(si.T)).to(dev)))
        so = torch.zeros(q.STEP_OUT, B, dtype=torch.float64, device=dev); sw = torch.zeros(q.STEP_STATE, B, dtype=torch.float64, device=dev)
        sd = torch.zeros(q.STEP_DIAG, B, dtype=torch.int32, device=dev)
        stream = torch.cuda.Stream(device=dev)
        def call(i):
            tk, st, si = sl[i % nrot]
            rc = mpc.lib.go1mpc_step_timing_step_batch(mpc.h, 3, B, tk.data_ptr(), st.data_ptr(), sw.data_ptr(), si.data_ptr(), so.data_ptr(), sd.data_ptr(), stream.cuda_stream)
            assert rc == 0
        for i in range(8): call(i)
        torch.cuda.synchronize()
        n = 60
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * n)]
        for i in range(n):
            ev[2 * i].record(stream); call(i); ev[2 * i + 1].record(stream); ev[2 * i + 1].synchronize()
        t = np.array([ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(n)])
        # back to back (throughput): 40 calls between two events
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(40): call(i)
        e1.record(stream); e1.synchronize()
        st_ = sd.cpu().numpy()[5::11][:3]
        print(f"{name} B={B} mode={os.environ.get('GO1MPC_STEP_MODE','auto')} lib={os.path.basename(os.environ.get('GO1MPC_LIB','default'))}: lone {1e3*t.mean():.1f} us (min {1e3*t.min():.1f}), back-to-back {1e3*e0.elapsed_time(e1)/40:.1f} us/call, infeasible {float((st_==2).mean()):.3f}", flush=True)
    mpc.close()
