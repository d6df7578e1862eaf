"""A/B of the body-MPC kernels: body_split (+ list-mode combined kernel) against body_fast alone
(GO1MPC_BODY_MODE=fast), same inputs: agreement of outputs / diagnostics, hand-over count, and
microseconds per launch at B = 4096 and 65536.  Run on the GPU box."""
import os, sys, subprocess
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    import quadrupedal_loco_b200 as q
    from quadrupedal_loco_b200 import synth
    nh = 10
    mpc = q.Go1Mpc(0); dev = torch.device("cuda", 0)
    stream = torch.cuda.ExternalStream(mpc.stream, device=dev)
    res = {}
    for B, scale in ((256, 1.0), (4096, 1.0), (65536, 1.0), (4096, 2.0)):
        d = synth.body_mpc_inputs(B, nh, seed=synth.SEED_CFG2 if scale == 1.0 else synth.SEED_CFG3, scale=scale)
        rec = q.pack_body_inputs(nh, d["tick"], d["tx"], d["theta"], d["bstate"], d["x_warm"], d["refs"])
        r = torch.from_numpy(rec).to(dev)
        o = torch.zeros(B, q.body_out_stride(nh), dtype=torch.float64, device=dev)
        dg = torch.zeros(B, q.body_diag_stride(nh), dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        h0 = mpc.body_handover_total()
        mpc.body_mpc_step(nh, B, r, o, dg); mpc.synchronize()
        h1 = mpc.body_handover_total()
        for _ in range(5): mpc.body_mpc_step(nh, B, r, o, dg)
        mpc.synchronize()
        n = 50 if B <= 4096 else 10
        with torch.cuda.stream(stream):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(n): mpc.body_mpc_step(nh, B, r, o, dg)
            e1.record(stream)
        e1.synchronize()
        us = e0.elapsed_time(e1) / n * 1e3
        res[f"{B}_{scale}"] = dict(out=o.cpu().numpy(), diag=dg.cpu().numpy())
        print(f"mode={os.environ.get('GO1MPC_BODY_MODE', 'tri')} B={B} scale={scale}: {us:.1f} us/launch, handed over {h1 - h0}, guard trips {mpc.body_guard_trips()}", flush=True)
    np.savez(sys.argv[2], **{k + "_out": v["out"] for k, v in res.items()}, **{k + "_diag": v["diag"] for k, v in res.items()})
    sys.exit(0)

os.makedirs("gpurun_out", exist_ok=True)
here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for var in os.environ.get("PROBE_VARIANTS", "").split():
    env = dict(os.environ, GO1MPC_BODY_MODE=os.environ.get("PROBE_VARIANT_MODE", "tri"), GO1MPC_LIB=os.path.join(here, "quadrupedal_loco_b200", "build", "variants", f"libgo1mpc_{var}.so"))
    print("variant", var, flush=True)
    subprocess.run([sys.executable, __file__, "child", f"/tmp/probe_var.npz"], env=env, check=True)
modes = os.environ.get("PROBE_MODES", "fast split tri").split()
for mode in modes:
    env = dict(os.environ, GO1MPC_BODY_MODE=mode)
    subprocess.run([sys.executable, __file__, "child", f"/tmp/probe_{mode}.npz"], env=env, check=True, timeout=150)
a = np.load("/tmp/probe_fast.npz")
for mode in modes[1:]:
  b = np.load(f"/tmp/probe_{mode}.npz")
  print("==== fast vs", mode)
  for k in a.files:
      if k.endswith("_diag"):
          da, db = a[k], b[k]
          same = (da[:, :9] == db[:, :9]).all(axis=1) & (da[:, 10:] == db[:, 10:]).all(axis=1)
          print(k, "diag rows equal:", int(same.sum()), "of", len(same), "| flops equal:", int((da[:, 9] == db[:, 9]).sum()),
                "| status!=0:", int((da[:, 0] > 0).sum()))
          bad = np.nonzero(~same)[0][:5]
          for i in bad: print("  row", i, "\n   fast ", da[i], "\n   split", db[i])
      else:
          oa, ob = a[k], b[k]
          err = np.abs(oa - ob) / np.maximum(1.0, np.abs(oa).max(axis=1, keepdims=True))
          print(k, "max rel diff", err.max(), "rows > 1e-9:", int((err.max(axis=1) > 1e-9).sum()))
