"""Three step-timing SQP launches at batch B (argv[1], default 65536): the target of an ncu capture."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quadrupedal_loco_b200 as q
from quadrupedal_loco_b200 import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
mpc = q.Go1Mpc(0); dev = torch.device("cuda", 0)
stick, sst, sinp = synth.step_timing_inputs(B, mpc.step_default_state(), seed=synth.SEED_CFG2)
st = torch.from_numpy(np.ascontiguousarray(sst.T)).to(dev); si = torch.from_numpy(np.ascontiguousarray(sinp.T)).to(dev)
tk = torch.from_numpy(stick.copy()).to(dev)
so = torch.zeros(q.STEP_OUT, B, dtype=torch.float64, device=dev); sto = st.clone()
sd = torch.zeros(q.STEP_DIAG, B, dtype=torch.int32, device=dev)
torch.cuda.synchronize()
for _ in range(3):
    rc = mpc.lib.go1mpc_step_timing_step_batch(mpc.h, 3, B, tk.data_ptr(), sto.data_ptr(), sto.data_ptr(), si.data_ptr(), so.data_ptr(), sd.data_ptr(), None)
    assert rc == 0
mpc.synchronize()
print("ok", int(sd[4].sum()))
