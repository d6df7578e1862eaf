"""GPU parity of the step-location / step-timing SQP tick (go1mpc_step_timing_step_batch,
through the C ABI) against the CPU oracle restatement of NLPClass::step_timing_opti_loop
(NLP/src/NLP/NLPClass_sqp.cpp:693-1102) and against outputs of the unmodified reference
(tests/golden/step_ref.npz): planner outputs and state to 1e-9 relative, identical active sets
and iteration counters for converged QPs, bit-exact integer step / phase indices."""
import numpy as np
import pytest

import quadrupedal_loco_b200 as q
from quadrupedal_loco_b200 import synth
from tests.test_oracle_vs_ref import load

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def gpu_tick(mpc, tick, state, inp, n_sqp=3, device=False):
    """Instance-major in, SoA through the ABI, instance-major out."""
    B = len(tick)
    t = np.ascontiguousarray(tick, np.int32)
    s = np.array(np.asarray(state).T, dtype=np.float64, order="C", copy=True)   # a (202, 1) transpose is already contiguous: force a copy
    i_ = np.array(np.asarray(inp).T, dtype=np.float64, order="C", copy=True)
    o = np.zeros((q.STEP_OUT, B)); d = np.full((q.STEP_DIAG, B), -7, np.int32)
    if device:
        import torch
        dev = torch.device("cuda", 0)
        tt, ts, ti, to, td = (torch.from_numpy(a).to(dev) for a in (t, s, i_, o, d))
        torch.cuda.synchronize()
        mpc.step_timing_step(n_sqp, B, tt, ts, ti, to, td)
        mpc.synchronize()
        s, o, d = ts.cpu().numpy(), to.cpu().numpy(), td.cpu().numpy()
    else:
        mpc.step_timing_step_host(n_sqp, B, t, s, i_, o, d)
    return o.T.copy(), s.T.copy(), d.T.copy()


def close(a, b, label):
    sc = np.maximum(1.0, np.abs(b))
    err = np.abs(a - b) / sc
    assert np.nanmax(err) < RTOL, f"{label}: rel err {np.nanmax(err):.3e} at {np.unravel_index(np.nanargmax(err), err.shape)}"


def assert_step_parity(go, gs, gd, oo, os_, od, label=""):
    # integer indices: bit-exact
    assert np.array_equal(gd[:, :5], od[:, :5]), f"{label}: period / phase indices"
    sts_g = gd[:, 5::11]; sts_o = od[:, 5::11]
    assert np.array_equal(sts_g, sts_o), f"{label}: QP status {np.nonzero((sts_g != sts_o).any(axis=1))[0][:8]}"
    # instances whose every QP converged: everything must agree
    ok = ((sts_o == 0) | (sts_o == -1)).all(axis=1) & np.isfinite(oo).all(axis=1)
    assert ok.sum() > 0
    for qi in range(5):
        o = 5 + 11 * qi
        conv = od[:, o] == 0
        assert np.array_equal(gd[conv, o + 1:o + 11], od[conv, o + 1:o + 11]), f"{label}: active set / counters of SQP iteration {qi}"
    close(go[ok], oo[ok], label + " out38")
    close(gs[ok], os_[ok], label + " state")
    return ok


@pytest.mark.parametrize("n_sqp", [1, 3, 5])
def test_step_parity_synth(mpc, oracle, n_sqp):
    B = 2048
    cfg = oracle.step_cfg(n_sqp)
    base = mpc.step_default_state()
    np.testing.assert_array_equal(base, oracle.step_default_state(cfg))
    tick, st, inp = synth.step_timing_inputs(B, base, seed=synth.SEED_CFG2 + n_sqp, amp=1.0)
    go, gs, gd = gpu_tick(mpc, tick, st, inp, n_sqp)
    os_ = st.copy()
    oo, od = oracle.step_tick_batch(cfg, tick, os_, inp)
    ok = assert_step_parity(go, gs, gd, oo, os_, od, f"K={n_sqp}")
    assert ok.mean() > 0.6
    assert (od[:, 6::11][:, :n_sqp][od[:, 5::11][:, :n_sqp] == 0] >= 3).any(), "no QP with several active constraints"


def test_step_device_entry_large_pushes(mpc, oracle):
    B = 4096
    cfg = oracle.step_cfg(3)
    tick, st, inp = synth.step_timing_inputs(B, mpc.step_default_state(), seed=synth.SEED_CFG3, amp=2.0)
    go, gs, gd = gpu_tick(mpc, tick, st, inp, 3, device=True)
    os_ = st.copy()
    oo, od = oracle.step_tick_batch(cfg, tick, os_, inp)
    assert_step_parity(go, gs, gd, oo, os_, od, "cfg3")
    assert (od[:, 5::11][:, :3] == 2).any()


def test_step_replay_closed_loop_vs_reference_golden(mpc):
    """cfg1: the 671-tick replay, state carried on the GPU side only, against the unmodified
    reference's outputs; integer indices bit-exact every tick."""
    g = load("step_ref.npz")
    T = g["replay_out"].shape[0] - 1
    st = g["replay_state"][1][None, :].copy()
    np.testing.assert_array_equal(mpc.step_default_state(), st[0])
    for i in range(1, T + 1):
        go, st, gd = gpu_tick(mpc, [i], st, g["replay_in"][i][None, :])
        assert list(gd[0, :4]) == list(g["replay_ints"][i]), (i, gd[0, :4], g["replay_ints"][i])
        close(go[0], g["replay_out"][i], f"replay out tick {i}")
        close(st[0], g["replay_state"][i + 1], f"replay state tick {i}")
        assert go[0, 27] == g["replay_out"][i][27] and go[0, 34] == g["replay_out"][i][34]


def test_step_pushes_vs_reference_golden(mpc):
    g = load("step_ref.npz")
    go, gs, gd = gpu_tick(mpc, g["push_tick"], g["push_state"], g["push_in"])
    fin = np.isfinite(g["push_out"]).all(axis=1)
    assert np.array_equal(gd[fin, :4], g["push_ints"][fin])
    conv = fin & (gd[:, 5::11][:, :3] == 0).all(axis=1)
    assert conv.sum() > 150
    close(go[conv], g["push_out"][conv], "push out")
    close(gs[conv], g["push_state_after"][conv], "push state")


def test_step_batch_of_one_and_ragged(mpc, oracle):
    cfg = oracle.step_cfg(3)
    for B in (1, 33, 127, 129):
        tick, st, inp = synth.step_timing_inputs(B, mpc.step_default_state(), seed=B, amp=0.5)
        go, gs, gd = gpu_tick(mpc, tick, st, inp)
        os_ = st.copy()
        oo, od = oracle.step_tick_batch(cfg, tick, os_, inp)
        assert_step_parity(go, gs, gd, oo, os_, od, f"B={B}")


def test_step_feedback_gains(mpc, oracle):
    lam = [0.25, 0.001, 0.025, 0.001]     # the commented multi-push preset, NLPClass_sqp.cpp:986-993
    h = q.Go1Mpc(0, cfg=dict(step=dict(lamda=lam)))
    try:
        cfg = oracle.step_cfg(3, lamda=lam)
        tick, st, inp = synth.step_timing_inputs(512, h.step_default_state(), seed=77, amp=0.7)
        rng = np.random.default_rng(5)
        inp[:, 0:6] = st[:, 189:195] + rng.uniform(-0.01, 0.01, (512, 6))
        go, gs, gd = gpu_tick(h, tick, st, inp)
        os_ = st.copy()
        oo, od = oracle.step_tick_batch(cfg, tick, os_, inp)
        assert_step_parity(go, gs, gd, oo, os_, od, "lamda")
    finally:
        h.close()


def test_step_full_size_properties(mpc):
    """65536 instances: idempotent launches; converged solves respect the step-period and
    reachability bounds; the period written back equals k_yu dt + log(tr1 + tr2) / w."""
    B = 65536
    tick, st, inp = synth.step_timing_inputs(B, mpc.step_default_state(), seed=synth.SEED_CFG3, amp=1.0)
    go, gs, gd = gpu_tick(mpc, tick, st, inp)
    go2, gs2, gd2 = gpu_tick(mpc, tick, st, inp)
    np.testing.assert_array_equal(go, go2); np.testing.assert_array_equal(gs, gs2); np.testing.assert_array_equal(gd, gd2)
    conv = (gd[:, 5::11][:, :3] == 0).all(axis=1)
    assert conv.mean() > 0.6
    v = gs[conv, 195:199]
    assert (v[:, 0] <= 0.15 + 1e-7).all() and (v[:, 0] >= -0.05 - 1e-7).all()
    Wn = np.sqrt(9.8 / 0.309458)
    k_yu = gd[conv, 1]
    ts_new = go[conv, 35]
    np.testing.assert_allclose(ts_new, k_yu * 0.025 + np.log(v[:, 2] + v[:, 3]) / Wn, rtol=1e-12)
    assert (ts_new <= 1.0 + 1e-6).all()


def test_step_out_of_place_state(mpc):
    """state_out_d != state_d: the input state stays untouched, the output equals the in-place result."""
    import torch
    B = 1000
    tick, st, inp = synth.step_timing_inputs(B, mpc.step_default_state(), seed=4, amp=1.0)
    go, gs, gd = gpu_tick(mpc, tick, st, inp)
    dev = torch.device("cuda", 0)
    tt = torch.from_numpy(np.ascontiguousarray(tick, np.int32)).to(dev)
    s_in = torch.from_numpy(np.array(st.T, order="C", copy=True)).to(dev); s_keep = s_in.clone()
    s_out = torch.full_like(s_in, float("nan"))
    ti = torch.from_numpy(np.array(inp.T, order="C", copy=True)).to(dev)
    to = torch.zeros(q.STEP_OUT, B, dtype=torch.float64, device=dev); td = torch.full((q.STEP_DIAG, B), -7, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    mpc.step_timing_step(3, B, tt, s_in, ti, to, td, state_out_d=s_out)
    mpc.synchronize()
    assert torch.equal(s_in, s_keep)
    np.testing.assert_array_equal(s_out.cpu().numpy().T, gs)
    np.testing.assert_array_equal(to.cpu().numpy().T, go)
    np.testing.assert_array_equal(td.cpu().numpy().T, gd)


def test_step_pipelined_host_entry_device_resident_state(mpc):
    """go1mpc_step_timing_step_batch_host_async: state stays on the device, only tick / inputs /
    results move; 5 closed-loop ticks equal the synchronous path."""
    import torch
    B = 777
    tick, st, inp = synth.step_timing_inputs(B, mpc.step_default_state(), seed=12, amp=0.6)
    dev = torch.device("cuda", 0)
    s_d = torch.from_numpy(np.array(st.T, order="C", copy=True)).to(dev)
    i_h = torch.from_numpy(np.array(inp.T, order="C", copy=True)).pin_memory()
    st_sync = st.copy()
    outs, ticks, diags = [], [], []
    for t in range(5):
        tk = torch.from_numpy((tick + t).astype(np.int32)).pin_memory()
        o = torch.zeros(q.STEP_OUT, B, dtype=torch.float64).pin_memory(); d = torch.zeros(q.STEP_DIAG, B, dtype=torch.int32).pin_memory()
        mpc.step_timing_step_host_async(3, B, tk.numpy(), s_d, i_h.numpy(), o.numpy(), d.numpy())
        outs.append(o); ticks.append(tk); diags.append(d)
    mpc.synchronize()
    for t in range(5):
        go, st_sync, gd = gpu_tick(mpc, tick + t, st_sync, inp)
        np.testing.assert_array_equal(outs[t].numpy().T, go, err_msg=f"tick +{t}")
        np.testing.assert_array_equal(diags[t].numpy().T, gd)
    np.testing.assert_array_equal(s_d.cpu().numpy().T, st_sync)


def test_foot_trajectory_replay_vs_reference_golden(mpc, oracle):
    """cfg1 on the GPU: the swing-foot tick of all 671 replay ticks against NLPClass's outputs.

    The reference's cubic fit is badly conditioned at some ticks (its first node, t_des - dt,
    comes within 1e-4 s of the mid-swing node): moving _ts by ONE ulp moves the reference's own
    accelerations by up to 6e-6 at 23 of the 671 ticks (measured with the oracle below).  So
    (a) tick by tick, from the reference's own state, the GPU must match to 1e-9 + 100 x that
    1-ulp sensitivity; (b) in closed loop (GPU planner state and GPU foot window carried for all
    671 ticks, where the planner's ts differs from the host's by libm ulps) right_support is
    bit-exact and positions stay within 2e-6 m."""
    g = load("step_ref.npz")
    cfg = oracle.step_cfg(3)
    T = g["replay_out"].shape[0] - 1
    sw0 = float(g["stepwidth0"][0])
    # (a) open loop, all ticks in one batch
    ticks = np.arange(1, T + 1, dtype=np.int32)
    after = g["replay_state"][2:T + 2]
    win = np.zeros((T, 32)); sens = np.zeros(T)
    fs = oracle.foot_default_state(sw0)[None, :].copy()
    for k, i in enumerate(ticks):
        win[k] = fs[0]
        f2 = fs.copy()
        a2 = after[k][None, :].copy(); a2[0, :27] = np.nextafter(a2[0, :27], np.inf)
        o2, _ = oracle.foot_tick_batch(cfg, [i], a2, [int(g["replay_out"][i][27])], f2, sw0)
        o1, _ = oracle.foot_tick_batch(cfg, [i], after[k][None, :], [int(g["replay_out"][i][27])], fs, sw0)
        sens[k] = np.abs(o2 - o1).max()
    o18 = np.zeros((18, T)); rs = np.zeros(T, np.int32)
    f = np.array(win.T, order="C", copy=True)
    mpc.foot_trajectory_host(T, ticks, np.array(after.T, order="C", copy=True), np.array(g["replay_out"][1:T + 1].T, order="C", copy=True), f, o18, rs)
    assert np.array_equal(rs, g["replay_right_support"][1:T + 1])
    err = np.abs(o18.T - g["replay_foot"][1:T + 1]).max(axis=1)
    assert (err <= 1e-9 + 100 * sens).all(), (err.max(), np.nonzero(err > 1e-9 + 100 * sens)[0][:5])
    assert (err[sens < 1e-12] < 1e-9).all()
    # (b) closed loop
    st = g["replay_state"][1][None, :].copy()
    fs = mpc.foot_default_state()[None, :].copy()
    for i in range(1, T + 1):
        go, st, gd = gpu_tick(mpc, [i], st, g["replay_in"][i][None, :])
        o1 = np.zeros((18, 1)); r1 = np.zeros(1, np.int32)
        f = np.array(fs.T, order="C", copy=True)
        mpc.foot_trajectory_host(1, np.array([i], np.int32), np.array(st.T, order="C", copy=True), np.array(go.T, order="C", copy=True), f, o1, r1)
        fs = f.T.copy()
        assert r1[0] == g["replay_right_support"][i], i
        assert np.abs(o1[:6, 0] - g["replay_foot"][i][:6]).max() < 2e-6, i


def test_foot_trajectory_batch_vs_oracle(mpc, oracle):
    """A batch of planners at different ticks of their walk (states from the replay), 3 consecutive ticks."""
    g = load("step_ref.npz")
    cfg = oracle.step_cfg(3)
    ticks0 = np.arange(40, 640, 3, dtype=np.int32); B = len(ticks0)
    st = g["replay_state"][ticks0].copy(); ost = st.copy()
    fs = np.tile(oracle.foot_default_state(), (B, 1))
    # a plausible window: feet on their nominal footholds of the running step
    rng = np.random.default_rng(3)
    fs[:, 0:24] += rng.uniform(-0.01, 0.01, (B, 24))
    ofs = fs.copy()
    for t in range(3):
        tick = ticks0 + t
        inp = g["replay_in"][tick]
        go, st, gd = gpu_tick(mpc, tick, st, inp)
        oo, od = oracle.step_tick_batch(cfg, tick, ost, inp)
        o18 = np.zeros((18, B)); rs = np.zeros(B, np.int32)
        f = np.array(fs.T, order="C", copy=True)
        mpc.foot_trajectory_host(B, tick, np.array(st.T, order="C", copy=True), np.array(go.T, order="C", copy=True), f, o18, rs)
        fs = f.T.copy()
        want, wrs = oracle.foot_tick_batch(cfg, tick, ost, oo[:, 27].astype(int), ofs)
        assert np.array_equal(rs, wrs)
        close(o18.T, want, f"foot out t+{t}")
        close(fs, ofs, f"foot window t+{t}")
