"""go1mpc_fused_tick_batch: one call = planner tick -> swing-foot trajectory -> body-MPC tick ->
servo kinematics with the body pose and the virtual feet taken from the ticks before it.  Every stage is pinned to
its oracle by its own test file; here the fused entry must reproduce, bit for bit, the four entry points called in
that order with the documented wiring (CoM = out38 rows 0..2, roll / pitch = body-MPC out[0], out[1], yaw 0,
right / left foot = out18 rows 0..2 / 3..5), and the joint angles it returns must place the feet on the targets
(CPU oracle forward kinematics, GO1/src/kinematics/Kinematics.cpp:145-229)."""
import numpy as np
import pytest

import quadrupedal_loco_b200 as q
from quadrupedal_loco_b200 import synth

pytestmark = pytest.mark.gpu


def make_inputs(mpc, B, seed, start=False):
    import torch
    dev = torch.device("cuda", 0)
    nh = 10
    tick, st, sin = synth.step_timing_inputs(B, mpc.step_default_state(), seed=seed)
    if start:
        # the first tick of the walk: planner state, swing-foot window and CoM are the consistent initial ones
        tick = np.ones(B, np.int32)
        st = np.tile(mpc.step_default_state(), (B, 1))
    foot0 = np.tile(mpc.foot_default_state(), (B, 1))
    d = synth.body_mpc_inputs(B, nh, seed=seed + 1)
    rec = q.pack_body_inputs(nh, d["tick"], d["tx"], d["theta"], d["bstate"], d["x_warm"], d["refs"])
    rng = np.random.Generator(np.random.Philox(seed + 2))
    # world-frame foot positions of the stand pose with the body 0.31 m above the ground (FR, FL, RR, RL)
    homing_leg = np.array([[0.1881, -0.12675, 0.0], [0.1881, 0.12675, 0.0], [-0.1881, -0.12675, 0.0], [-0.1881, 0.12675, 0.0]])
    homing = np.tile(homing_leg.reshape(1, 12), (B, 1)) + rng.uniform(-0.01, 0.01, (B, 12))
    q0 = np.tile(np.array([0.0, 0.67, -1.3] * 4), (B, 1)) + rng.uniform(-0.1, 0.1, (B, 12))
    soa = lambda a: torch.from_numpy(np.array(np.asarray(a, dtype=np.float64).T, order="C", copy=True)).to(dev)
    return dict(nh=nh, B=B,
                tick=torch.from_numpy(np.ascontiguousarray(tick, np.int32)).to(dev), state=soa(st), sin=soa(sin), foot=soa(foot0),
                rec=torch.from_numpy(rec).to(dev), homing=soa(homing), q=soa(q0))


def run_separate(mpc, I, gait_mode, y_offset):
    import torch
    dev = I["rec"].device; B, nh = I["B"], I["nh"]
    f64 = dict(dtype=torch.float64, device=dev); i32 = dict(dtype=torch.int32, device=dev)
    st = I["state"].clone(); foot = I["foot"].clone(); qq = I["q"].clone()
    out38 = torch.zeros(q.STEP_OUT, B, **f64); sdiag = torch.zeros(q.STEP_DIAG, B, **i32)
    out18 = torch.zeros(18, B, **f64); rs = torch.zeros(B, **i32)
    bout = torch.zeros(B, q.body_out_stride(nh), **f64); bdiag = torch.zeros(B, q.body_diag_stride(nh), **i32)
    jac = torch.zeros(36, B, **f64); fdes = torch.zeros(12, B, **f64); its = torch.zeros(4, B, **i32)
    mpc.step_timing_step(3, B, I["tick"], st, I["sin"], out38, sdiag)
    mpc.foot_trajectory(B, I["tick"], st, out38, foot, out18, rs)
    mpc.body_mpc_step(nh, B, I["rec"], bout, bdiag)
    mpc.synchronize()
    theta = torch.zeros(3, B, **f64)
    theta[0] = bout[:, 0]; theta[1] = bout[:, 1]
    com = out38[0:3].contiguous(); rfoot = out18[0:3].contiguous(); lfoot = out18[3:6].contiguous()
    torch.cuda.synchronize()
    mpc.servo_kin_tick(B, gait_mode, y_offset, com, theta, rfoot, lfoot, I["homing"], qq, jac, fdes, its)
    mpc.synchronize()
    return dict(state=st, foot=foot, q=qq, out38=out38, sdiag=sdiag, out18=out18, rs=rs, bout=bout, bdiag=bdiag, jac=jac,
                fdes=fdes, its=its, theta=theta)


def run_fused(mpc, I, gait_mode, y_offset):
    import torch
    dev = I["rec"].device; B, nh = I["B"], I["nh"]
    f64 = dict(dtype=torch.float64, device=dev); i32 = dict(dtype=torch.int32, device=dev)
    st = I["state"].clone(); foot = I["foot"].clone(); qq = I["q"].clone()
    out38 = torch.zeros(q.STEP_OUT, B, **f64); sdiag = torch.zeros(q.STEP_DIAG, B, **i32)
    out18 = torch.zeros(18, B, **f64); rs = torch.zeros(B, **i32)
    bout = torch.zeros(B, q.body_out_stride(nh), **f64); bdiag = torch.zeros(B, q.body_diag_stride(nh), **i32)
    jac = torch.zeros(36, B, **f64); fdes = torch.zeros(12, B, **f64); its = torch.zeros(4, B, **i32)
    theta = torch.zeros(3, B, **f64)
    torch.cuda.synchronize()
    l0 = mpc.launch_count
    mpc.fused_tick(B, 3, I["tick"], st, I["sin"], out38, foot, out18, nh, I["rec"], bout, gait_mode, y_offset, I["homing"], qq,
                   theta, step_diag=sdiag, right_support=rs, body_diag=bdiag, jac=jac, foot_des=fdes, ik_iters=its)
    mpc.synchronize()
    return dict(state=st, foot=foot, q=qq, out38=out38, sdiag=sdiag, out18=out18, rs=rs, bout=bout, bdiag=bdiag, jac=jac,
                fdes=fdes, its=its, theta=theta), mpc.launch_count - l0


@pytest.mark.parametrize("B,gait_mode", [(1, 102), (257, 101), (3000, 102), (4096, 103)])
def test_fused_tick_equals_the_four_entry_points(mpc, B, gait_mode):
    I = make_inputs(mpc, B, seed=41 + B)
    a = run_separate(mpc, I, gait_mode, 0.7)
    b, launches = run_fused(mpc, I, gait_mode, 0.7)
    assert launches >= 5
    for k in a:
        x, y = a[k].cpu().numpy(), b[k].cpu().numpy()
        assert np.array_equal(x, y, equal_nan=True), f"{k} differs between the fused tick and the separate entry points"
    assert (b["bdiag"][:, 0].cpu().numpy() <= 0).mean() > 0.5


def test_fused_tick_feet_reach_their_targets(mpc, oracle):
    """The servo stage of the fused tick against the CPU oracle: forward kinematics of the returned joint angles,
    with the body pose the fused tick used, lands on the foot targets it reports (IK stops at |dp|^2 <= 1e-6)."""
    B = 64
    I = make_inputs(mpc, B, seed=7, start=True)
    r, _ = run_fused(mpc, I, 102, 0.7)
    qj = r["q"].cpu().numpy().T.reshape(B, 4, 3); fdes = r["fdes"].cpu().numpy().T.reshape(B, 4, 3)
    com = r["out38"].cpu().numpy()[0:3].T.copy(); com[:, 1] *= 0.7
    th = r["theta"].cpu().numpy().T.copy()
    its = r["its"].cpu().numpy().T
    assert (its <= 15).all()
    for leg in range(4):
        pos, _ = oracle.leg_fk(qj[:, leg], np.full(B, leg, np.int32), com, th)
        conv = its[:, leg] < 15
        assert conv.mean() > 0.9
        assert np.abs(pos[conv] - fdes[conv, leg]).max() < 2e-3


def test_fused_tick_grf_stage_equals_separate_calls(mpc):
    """Optional stage 5: the force QP and the torque map on the Jacobians of the servo stage, bit-identical to
    go1mpc_grf_force_opt_batch + go1mpc_grf_joint_torques_batch called after the four-stage tick."""
    import torch
    from tests.test_grf import grf_inputs, tau_inputs
    B = 1500
    I = make_inputs(mpc, B, seed=91)
    dev = I["rec"].device
    f64 = dict(dtype=torch.float64, device=dev); i32 = dict(dtype=torch.int32, device=dev)
    a = run_separate(mpc, I, 102, 0.7)
    d = grf_inputs(B, seed=92)
    rec = np.zeros((B, 48))
    rec[:, 0:3] = d["base"]; rec[:, 3:15] = d["legs"]; rec[:, 15:21] = d["FT"]; rec[:, 21:33] = d["prev"] * 0.5; rec[:, 33:45] = d["prev"]
    rec[:, 45] = d["mode"]; rec[:, 46] = d["rs"]
    t = tau_inputs(B, seed=93)
    soa = lambda x: torch.from_numpy(np.array(x.reshape(B, -1).T, order="C", copy=True)).to(dev)
    gin = torch.from_numpy(rec).to(dev)
    sw, pd, pe, vd, ve = (soa(t[k]) for k in ("swing", "p_des", "p_est", "pv_des", "pv_est"))
    gout_a = torch.zeros(B, 16, **f64); gdg_a = torch.zeros(B, 32, **i32); tau_a = torch.zeros(12, B, **f64)
    torch.cuda.synchronize()
    mpc.grf_force_opt(B, gin, gout_a, gdg_a)
    mpc.grf_joint_torques(B, a["jac"], sw, pd, pe, vd, ve, gout_a, tau_a, F_strides=(1, 16))
    mpc.synchronize()
    # fused
    st = I["state"].clone(); foot = I["foot"].clone(); qq = I["q"].clone()
    out38 = torch.zeros(q.STEP_OUT, B, **f64); out18 = torch.zeros(18, B, **f64)
    bout = torch.zeros(B, q.body_out_stride(I["nh"]), **f64); jac = torch.zeros(36, B, **f64); theta = torch.zeros(3, B, **f64)
    gout_b = torch.zeros(B, 16, **f64); gdg_b = torch.zeros(B, 32, **i32); tau_b = torch.full((12, B), np.nan, **f64)
    torch.cuda.synchronize()
    mpc.fused_tick(B, 3, I["tick"], st, I["sin"], out38, foot, out18, I["nh"], I["rec"], bout, 102, 0.7, I["homing"], qq, theta, jac=jac,
                   grf_in=gin, grf_out=gout_b, grf_diag=gdg_b, swing=sw, p_des=pd, p_est=pe, pv_des=vd, pv_est=ve, tau=tau_b)
    mpc.synchronize()
    same = lambda x, y: np.array_equal(x.cpu().numpy(), y.cpu().numpy(), equal_nan=True)   # an infeasible planner tick leaves NaN poses
    assert same(jac, a["jac"]) and same(qq, a["q"])
    assert same(gout_a, gout_b) and same(gdg_a, gdg_b)
    assert same(tau_a, tau_b)
    ok = torch.isfinite(jac).all(dim=0)
    assert ok.float().mean() > 0.5 and torch.isfinite(tau_b[:, ok]).all()
    assert (gdg_b[:, 0] == 0).float().mean() > 0.5
