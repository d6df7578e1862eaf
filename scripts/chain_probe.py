"""Throughput of the two node entries on one GPU: go1mpc_nlp_node_tick_batch (40 Hz planner node) and
go1mpc_rt_node_tick_batch (100 Hz node, nh = 10) for B robots walking out of phase, device-resident, CUDA events."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quadrupedal_loco_b200 as q
mpc = q.Go1Mpc(0); dev = torch.device("cuda", 0)
stream = torch.cuda.ExternalStream(mpc.stream, device=dev)
nh = 10
for B in (4096, 65536):
    f64 = dict(dtype=torch.float64, device=dev)
    nlp_st = torch.from_numpy(np.repeat(mpc.nlp_node_default_state()[:, None], B, axis=1).copy()).to(dev)
    rt_st = torch.from_numpy(np.repeat(mpc.rt_node_default_state(nh)[:, None], B, axis=1).copy()).to(dev)
    body_in = torch.zeros(B, q.body_in_stride(nh), **f64); body_out = torch.zeros(B, q.body_out_stride(nh), **f64)
    gait = torch.zeros(100, B, **f64); traj = torch.zeros(100, B, **f64)
    rng = np.random.default_rng(3)
    phase = torch.from_numpy(rng.integers(0, 30, B).astype(np.int32)).to(dev)       # robots start up to 30 slow ticks apart
    rf = torch.from_numpy(rng.uniform(-0.01, 0.01, (3, B))).to(dev); lf = torch.from_numpy(rng.uniform(-0.01, 0.01, (3, B))).to(dev)
    wds = [(torch.full((B,), c, dtype=torch.int32, device=dev) - phase).clamp(min=0).contiguous() for c in range(0, 330)]
    torch.cuda.synchronize()
    # walk every robot well into the gait (slow ticks 1..249, fast ticks alongside), then time 60 slow + 150 fast ticks
    count, t_ms = 0, 0
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t_slow = t_fast = 0.0; n_slow = n_fast = 0
    while count < 320 or t_ms % 25:
        timed = count >= 250
        if t_ms % 25 == 0:
            count += 1
            if timed: ev[0].record(stream)
            mpc.nlp_node_tick(B, nlp_st, wds[count], gait, rfoot_fb_d=rf, lfoot_fb_d=lf)
            if timed: ev[1].record(stream); ev[1].synchronize(); t_slow += ev[0].elapsed_time(ev[1]); n_slow += 1
        if t_ms % 10 == 0:
            if timed: ev[2].record(stream)
            mpc.rt_node_tick(nh, B, rt_st, gait, body_in, body_out, traj)
            if timed: ev[3].record(stream); ev[3].synchronize(); t_fast += ev[2].elapsed_time(ev[3]); n_fast += 1
        t_ms += 5
    mpc.synchronize()
    ok = torch.isfinite(traj).all().item() and torch.isfinite(gait).all().item()
    live = (traj[72:86].abs().sum(dim=0) > 0).float().mean().item()
    print(f"B={B:6d}: planner node {t_slow / n_slow * 1e3:8.1f} us per tick ({B / (t_slow / n_slow) / 1e3:.1f} M robot ticks/s, 3 QP solves each), "
          f"100 Hz node {t_fast / n_fast * 1e3:8.1f} us per tick ({B / (t_fast / n_fast) / 1e3:.1f} M robot ticks/s, 1 QP solve each); "
          f"finite {ok}, robots with a live body MPC {live:.2f}")
