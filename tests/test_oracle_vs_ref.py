"""Pins the CPU oracle against the reference itself.

tests/golden/*.npz hold outputs of the UNMODIFIED reference sources (EiQuadProg.cpp,
QPBaseClass.cpp, PRMPCClass.cpp, Kinematics.cpp) compiled against oracle/eigen_shim
(tests/golden/make_golden.py, run in the authoring container).  The oracle restatement must
reproduce them BIT FOR BIT: that pins control flow, tie-breaking, tolerances and every
reference quirk.  (It cannot pin real Eigen's summation order: Eigen is absent here.)
Where oracle/_ref is present the same comparison is also made live on fresh inputs."""
import ctypes
import os

import numpy as np
import pytest

from quadrupedal_loco_b200 import synth
from tests.oracle_lib import P, PI, ref_path

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name))


def qp_cases():
    g = load("qp_ref.npz")
    ci = 0
    while f"c{ci}_shape" in g:
        n, p, m, B, seed = (int(v) for v in g[f"c{ci}_shape"])
        yield ci, n, p, m, B, {k: g[f"c{ci}_{k}"] for k in ("G", "g0", "CE", "ce0", "CI", "ci0", "x", "cost", "active", "nactive")}
        ci += 1


def test_golden_inputs_regenerate():
    """The stored inputs are what synth produces from the stored seeds (generators did not drift)."""
    from tests.golden.make_golden import QP_CASES
    g = load("qp_ref.npz")
    for ci, (n, p, m, B, seed, kw) in enumerate(QP_CASES):
        d = synth.random_qp(B, n, p, m, seed=seed, **kw)
        for k in ("G", "g0", "CI", "ci0"):
            np.testing.assert_array_equal(d[k], g[f"c{ci}_{k}"])


def test_oracle_qp_bit_exact_vs_reference_golden(oracle):
    ncase = 0
    for ci, n, p, m, B, d in qp_cases():
        o = oracle.qp_solve_batch(n, p, m, d)
        for b in range(B):
            if np.isinf(d["cost"][b]):
                assert np.isinf(o["cost"][b]) and o["status"][b] in (1, 2)
            else:
                assert o["cost"][b] == d["cost"][b], (ci, b)
                k = d["nactive"][b]
                assert o["nactive"][b] == k and np.array_equal(o["active"][b, :k], d["active"][b, :k]), (ci, b)
            np.testing.assert_array_equal(o["x"][b], d["x"][b])   # also for infeasible exits: same garbage
        ncase += 1
    assert ncase >= 6


def test_oracle_body_bit_exact_vs_reference_golden(oracle):
    """40 closed-loop ticks of 48 instances at the reference's compile-time horizon nh = 4."""
    g = load("body_ref_nh4.npz")
    nh = 4
    T, B = g["out14"].shape[:2]
    cfg = oracle.body_cfg(nh)
    theta = g["theta0"].copy(); x = np.zeros((B, 2 * nh)); o14 = np.zeros((B, 14))
    for t in range(T):
        oracle.body_step_batch(cfg, g["tick0"] + t, g["tx"], theta, g["bstate"], g["refs"], o14, x)
        np.testing.assert_array_equal(o14, g["out14"][t], err_msg=f"out14 tick {t}")
        np.testing.assert_array_equal(theta, g["theta"][t], err_msg=f"state tick {t}")
        np.testing.assert_array_equal(x, g["vini"][t], err_msg=f"V_ini tick {t}")
    assert (np.abs(g["out14"]).sum(axis=2) > 0).mean() > 0.9


@pytest.mark.skipif(ref_path("libref.so") is None, reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_qp_vs_live_reference(oracle):
    lib = ctypes.CDLL(ref_path("libref.so"))
    for (n, p, m, seed, kw) in [(4, 1, 24, 71, dict(paired=True)), (10, 0, 30, 72, dict(dup=True)), (16, 4, 40, 73, dict(paired=True)),
                                (7, 0, 12, 74, dict(infeasible_frac=0.5))]:
        d = synth.random_qp(40, n, p, m, seed=seed, **kw)
        o = oracle.qp_solve_batch(n, p, m, d)
        for b in range(40):
            x = np.zeros(n); c = np.zeros(1); a = np.zeros(m + p + 1, np.int32); na = np.zeros(1, np.int32)
            lib.ref_qp_solve(n, p, m, P(d["G"][b].copy()), P(d["g0"][b].copy()), P(d["CE"][b].copy()), P(d["ce0"][b].copy()),
                             P(d["CI"][b].copy()), P(d["ci0"][b].copy()), P(x), P(c), PI(a), PI(na))
            np.testing.assert_array_equal(o["x"][b], x)
            assert (np.isinf(c[0]) and np.isinf(o["cost"][b])) or c[0] == o["cost"][b]
            if not np.isinf(c[0]):
                assert na[0] == o["nactive"][b] and np.array_equal(a[:na[0]], o["active"][b, :na[0]])
            # through the reference's own container: QPBaseClass::solveQP's bool == "no NaN in X"
            x2 = np.zeros(n)
            ok = lib.ref_qpbase_solve(n, p, m, P(d["G"][b].copy()), P(d["g0"][b].copy()), P(d["CE"][b].copy()),
                                      P(d["ce0"][b].copy()), P(d["CI"][b].copy()), P(d["ci0"][b].copy()), P(x2))
            assert ok == 1
            np.testing.assert_array_equal(x2, x)


def _step_cfg_from_golden(oracle, g, n_sqp=3):
    names = ("dt Wn ggg t_min t_max footx_max footx_min footx_vmax footx_vmin footy_vmax footy_vmin comax_max comax_min "
             "comay_max comay_min aax aay aaxv aayv bbx bby rr1 rr2 half_hip_width foot_width").split()
    cfg = oracle.step_cfg(n_sqp)
    for n, v in zip(names, g["consts"]):
        assert getattr(cfg, n) == v, f"default {n} differs from the reference's constant"
    assert cfg.hcom == g["consts"][29]
    return cfg


def test_oracle_step_replay_bit_exact_vs_reference_golden(oracle):
    """cfg1: the deterministic 671-tick replay of NLPClass::step_timing_opti_loop -- outputs,
    updated state and integer indices, bit for bit, every tick."""
    g = load("step_ref.npz")
    cfg = _step_cfg_from_golden(oracle, g)
    T = g["replay_out"].shape[0] - 1
    # default tables are the reference's Initialize()
    np.testing.assert_array_equal(oracle.step_default_state(cfg), g["replay_state"][1])
    st = g["replay_state"][1].copy()
    for i in range(1, T + 1):
        np.testing.assert_array_equal(st, g["replay_state"][i], err_msg=f"state before tick {i}")
        out, dg = oracle.step_tick_batch(cfg, [i], st[None, :], g["replay_in"][i][None, :])
        # step_tick_batch updates a copy of the row when given a fresh 2-D view: redo on a real 2-D array
        s2 = g["replay_state"][i].copy()[None, :]
        out, dg = oracle.step_tick_batch(cfg, [i], s2, g["replay_in"][i][None, :])
        st = s2[0]
        np.testing.assert_array_equal(out[0], g["replay_out"][i], err_msg=f"out38 tick {i}")
        assert list(dg[0, :4]) == list(g["replay_ints"][i]), (i, dg[0, :4], g["replay_ints"][i])
    np.testing.assert_array_equal(st, g["replay_state"][T + 1])


def test_oracle_step_pushes_bit_exact_vs_reference_golden(oracle):
    g = load("step_ref.npz")
    cfg = _step_cfg_from_golden(oracle, g)
    st = g["push_state"].copy()
    out, dg = oracle.step_tick_batch(cfg, g["push_tick"], st, g["push_in"])
    fin = np.isfinite(g["push_out"]).all(axis=1)
    assert fin.sum() > 300
    np.testing.assert_array_equal(out[fin], g["push_out"][fin])
    np.testing.assert_array_equal(st[fin], g["push_state_after"][fin])
    np.testing.assert_array_equal(dg[fin, :4], g["push_ints"][fin])
    sts = dg[:, 5::11][:, :3]
    assert (sts == 0).any() and (sts == 2).any()       # both converged and infeasible solves are pinned
    assert (dg[:, 6::11][:, :3][sts == 0] >= 3).any()   # with several constraints active


@pytest.mark.skipif(ref_path("libref_nlp.so") is None, reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_step_vs_live_reference(oracle):
    lib = ctypes.CDLL(ref_path("libref_nlp.so"))
    lib.ref_nlp_new.restype = ctypes.c_void_p
    lib.ref_nlp_new.argtypes = [ctypes.c_double] * 3
    h = ctypes.c_void_p(lib.ref_nlp_new(0.06, 0.2535, 0.0))     # a different step length than the golden run
    cfg = oracle.step_cfg(3)
    est = np.zeros(18); rf = np.array([0, -0.12675, 0.]); lf = np.array([0, 0.12675, 0.])
    for i in range(1, 200):
        st = np.zeros(202); lib.ref_nlp_get_state(h, i, P(st))
        out = np.zeros(38); hz = np.zeros(10); ints = np.zeros(4, np.int32)
        lib.ref_nlp_step(h, i, P(est), P(rf), P(lf), 0, P(out), P(hz), PI(ints))
        after = np.zeros(202); lib.ref_nlp_get_state(h, i + 1, P(after))
        inp = np.zeros(20); inp[6:8] = rf[:2]; inp[8:10] = lf[:2]; inp[10:13] = hz[:3]; inp[13:16] = hz[3:6]; inp[16:19] = hz[6:9]; inp[19] = hz[9]
        s2 = st[None, :].copy()
        o, dg = oracle.step_tick_batch(cfg, [i], s2, inp[None, :])
        np.testing.assert_array_equal(o[0], out); np.testing.assert_array_equal(s2[0], after)
        assert list(dg[0, :4]) == list(ints)
    lib.ref_nlp_free(h)


def test_oracle_foot_trajectory_bit_exact_vs_reference_golden(oracle):
    """NLPClass::Foot_trajectory_solve_mod2 over the 671-tick replay: the oracle carries only its
    32-double window of the reference's whole-walk foot arrays and must reproduce the Vec18 and
    right_support of every tick bit for bit."""
    g = load("step_ref.npz")
    cfg = _step_cfg_from_golden(oracle, g)
    T = g["replay_out"].shape[0] - 1
    sw0 = float(g["stepwidth0"][0])
    fs = oracle.foot_default_state(sw0)[None, :].copy()
    for i in range(1, T + 1):
        after = g["replay_state"][i + 1][None, :]
        out, rs = oracle.foot_tick_batch(cfg, [i], after, [int(g["replay_out"][i][27])], fs, sw0)
        np.testing.assert_array_equal(out[0], g["replay_foot"][i], err_msg=f"tick {i}")
        assert rs[0] == g["replay_right_support"][i], i
    assert set(np.unique(g["replay_right_support"][1:])) == {0, 1, 2}


def _oracle_interp(oracle, d, nh, t_end, dt=0.025, dt_sample=0.01):
    lib = oracle.lib
    lib.orc_interp_aaa_inv_mod.argtypes = [ctypes.c_double, ctypes.c_void_p]; lib.orc_interp_aaa_inv_mod.restype = None
    lib.orc_interp_position_mod3.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double] + [ctypes.c_void_p] * 5
    lib.orc_interp_position_mod3.restype = None
    inv = np.zeros(16); lib.orc_interp_aaa_inv_mod(dt, P(inv))
    N = len(d["walktime"]); out = np.full((N, 9 + 3 * (nh - 1)), np.nan)
    for b in range(N):
        s = d["samples"][b]
        lib.orc_interp_position_mod3(P(inv), nh, t_end, int(d["walktime"][b]), dt_sample, P(s[0].copy()), P(s[1].copy()), P(s[2].copy()),
                                     P(s[3].copy()), P(out[b]))
    return inv, out


def test_oracle_ref_interp_bit_exact_vs_reference_golden(oracle):
    """oracle/ref_interp.c against PRMPCClass::XGetSolution_position_mod3 (golden vectors of the unmodified class at
    its nh = 4): the 4x4 inverse and every interpolated value bit for bit, zeros beyond _t_end_footstep; at the
    sample instants the cubic returns the samples (interpolation property, any horizon)."""
    g = load("interp_ref.npz")
    nh, t_end = int(g["nh"][0]), int(g["t_end"][0])
    d = {k: g[k] for k in ("samples", "walktime")}
    inv, out = _oracle_interp(oracle, d, nh, t_end)
    np.testing.assert_array_equal(inv, g["inv"])
    np.testing.assert_array_equal(out, g["out"][:, :9 + 3 * (nh - 1)])
    assert (g["out"][:, 9 + 3 * (nh - 1):] == 0).all()
    late = d["walktime"] > t_end
    assert late.any() and (out[late] == 0).all() and (np.abs(out[~late]).sum(axis=1) > 0).all()
    # the cubic passes through its four samples: walktime 0 with dt_sample = dt evaluates t = 0, dt, 2 dt
    one = dict(samples=d["samples"][:50], walktime=np.zeros(50, np.int32))
    _, o3 = _oracle_interp(oracle, one, 3, t_end, dt_sample=0.025)
    s = one["samples"]
    np.testing.assert_allclose(o3[:, 0:3], s[:, 1], atol=1e-12)
    np.testing.assert_allclose(o3[:, 9:12], s[:, 2], atol=1e-12)
    np.testing.assert_allclose(o3[:, 12:15], s[:, 3], atol=1e-12)


@pytest.mark.skipif(ref_path("libref_rt.so") is None, reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_ref_interp_vs_live_reference(oracle):
    from tests.golden.make_golden import interp_inputs
    rtl = ctypes.CDLL(ref_path("libref_rt.so"))
    if not hasattr(rtl, "ref_body_position_mod3"):
        pytest.skip("oracle/_ref predates ref_body_position_mod3")
    rtl.ref_body_new.restype = ctypes.c_void_p
    h = ctypes.c_void_p(rtl.ref_body_new())
    d = interp_inputs(400, seed=141)
    want = np.zeros((400, 21)); inv = np.zeros(16)
    for b in range(400):
        s = d["samples"][b]
        t_end = rtl.ref_body_position_mod3(h, int(d["walktime"][b]), ctypes.c_double(0.01), P(s[0].copy()), P(s[1].copy()), P(s[2].copy()),
                                           P(s[3].copy()), P(want[b]), P(inv))
    nh = rtl.ref_body_nh()
    rtl.ref_body_free(h)
    oinv, out = _oracle_interp(oracle, d, nh, t_end)
    np.testing.assert_array_equal(oinv, inv)
    np.testing.assert_array_equal(out, want[:, :9 + 3 * (nh - 1)])


class _FootRotState(ctypes.Structure):
    _fields_ = [("bjxx", ctypes.c_int), ("bjx1", ctypes.c_int), ("Rr", ctypes.c_double * 15), ("Lr", ctypes.c_double * 15)]


def _oracle_foot_rot(oracle, tx, ts, td, footx, scal, nh, ticks, dt_sample=0.01):
    lib = oracle.lib
    lib.orc_foot_rotation.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                                             ctypes.c_int, ctypes.c_double, ctypes.c_void_p]
    lib.orc_foot_rotation.restype = None
    s = _FootRotState(); lib.orc_foot_rot_state_init(ctypes.byref(s))
    out = np.full((len(ticks), 6 * nh), np.nan)
    for k, t in enumerate(ticks):
        lib.orc_foot_rotation(P(tx), P(ts), P(td), P(footx), float(scal[0]), float(scal[1]), int(scal[2]), nh, ctypes.byref(s), int(t), dt_sample, P(out[k]))
    return out


def test_oracle_foot_rotation_bit_exact_vs_reference_golden(oracle):
    """oracle/foot_rot.c against PRMPCClass::XGetSolution_Foot_rotation: a 574-call sequence on one object (angle members
    persist), forward / backward / zero-length steps, ticks beyond _t_end_footstep and jumps back -- bit for bit."""
    g = load("foot_rot_ref.npz")
    nh = int(g["nh"][0])
    out = _oracle_foot_rot(oracle, g["tx"].copy(), g["ts"].copy(), g["td"].copy(), g["footx"].copy(), g["scal"], nh, g["ticks"])
    np.testing.assert_array_equal(out, g["out"][:, :6 * nh])
    assert (g["out"][:, 6 * nh:] == 0).all()
    o = out.reshape(len(out), nh, 6)
    assert (o[:, :, 0] <= 0).all() and (o[:, :, 3] >= 0).all()          # right roll bumps down, left roll up
    assert (o[:, :, 1] <= 0).all() and (o[:, :, 4] <= 0).all() and (o[:, :, [1, 4]] < 0).any()   # pitch only toes-down
    assert (o[:, :, [2, 5]] == 0).all()                                  # yaw is never written
    assert np.abs(o[:, :, 0]).max() <= 0.13 + 1e-12 and np.abs(o[:, :, 3]).max() <= 0.15 + 1e-12


@pytest.mark.skipif(ref_path("libref_rt.so") is None, reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_foot_rotation_vs_live_reference(oracle):
    from tests.golden.make_golden import foot_rot_inputs
    rtl = ctypes.CDLL(ref_path("libref_rt.so"))
    if not hasattr(rtl, "ref_body_foot_rotation"):
        pytest.skip("oracle/_ref predates ref_body_foot_rotation")
    rtl.ref_body_new.restype = ctypes.c_void_p
    h = ctypes.c_void_p(rtl.ref_body_new())
    tx = np.zeros(27); ts = np.zeros(27); td = np.zeros(27); fx0 = np.zeros(27); sc = np.zeros(4)
    rtl.ref_body_foot_tables(h, P(tx), P(ts), P(td), P(fx0), P(sc))
    footx, ticks = foot_rot_inputs(seed=151)
    ticks = ticks[::2]
    rtl.ref_body_set_footx(h, P(footx.copy()))
    want = np.zeros((len(ticks), 30))
    for k, t in enumerate(ticks):
        rtl.ref_body_foot_rotation(h, int(t), ctypes.c_double(0.01), P(want[k]))
    nh = rtl.ref_body_nh()
    rtl.ref_body_free(h)
    out = _oracle_foot_rot(oracle, tx, ts, td, footx, sc, nh, ticks)
    np.testing.assert_array_equal(out, want[:, :6 * nh])


class _RtFootCfg(ctypes.Structure):
    _fields_ = [(n, ctypes.c_double) for n in "dt dt_mpc tstep tdsp_ratio stepwidth0 lift_height".split()]


def oracle_rt_foot_run(oracle, nh, nrt, stop, ticks=None):
    """oracle/rt_foot.c over a call sequence on one state; returns (out [T+1, 6 (nh+1)], state after the last call)."""
    lib = oracle.lib
    lib.orc_rt_foot_state_doubles.restype = ctypes.c_int
    c = _RtFootCfg(); lib.orc_rt_foot_cfg_default(ctypes.byref(c))
    S = lib.orc_rt_foot_state_doubles(nh)
    s = np.zeros(S); lib.orc_rt_foot_state_default(ctypes.byref(c), nh, P(s))
    s0 = s.copy()
    T = len(nrt) - 1
    out = np.zeros((T + 1, 6 * (nh + 1)))
    for j in (range(1, T + 1) if ticks is None else ticks):
        lib.orc_rt_foot_traj(ctypes.byref(c), nh, P(s), int(j), int(stop[j]), P(nrt[j].copy()), P(out[j]))
    return out, s, s0


def test_oracle_rt_foot_bit_exact_vs_reference_golden(oracle):
    """oracle/rt_foot.c against PRMPCClass::Foot_trajectory_solve_mod2 (RT/src/FastMPC/PRMPCClass.cpp:1756-2195): the 1750-call
    100 Hz sequence of tests/golden/rt_foot_ref.npz on one object -- every returned foot position, the initial members and the
    members left after the last call, bit for bit."""
    g = load("rt_foot_ref.npz")
    nh = int(g["nh"][0])
    out, s_end, s0 = oracle_rt_foot_run(oracle, nh, g["nrt"], g["stop"])
    np.testing.assert_array_equal(s0, g["state0"])
    np.testing.assert_array_equal(out[1:], g["out"][1:, :6 * (nh + 1)])
    np.testing.assert_array_equal(s_end, g["state_end"])
    assert (np.abs(np.diff(g["out"][:, 2])) > 0).sum() > 300          # the swing foot is lifted: cubic branch exercised


@pytest.mark.skipif(ref_path("libref_rt.so") is None, reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_rt_foot_vs_live_reference(oracle):
    from tests.golden.make_golden import rt_foot_inputs
    rtl = ctypes.CDLL(ref_path("libref_rt.so"))
    if not hasattr(rtl, "ref_body_foot_traj"):
        pytest.skip("oracle/_ref predates ref_body_foot_traj")
    nrt, stop = rt_foot_inputs(seed=161)
    rtl.ref_body_new.restype = ctypes.c_void_p
    h = ctypes.c_void_p(rtl.ref_body_new())
    nh = rtl.ref_body_nh()
    T = len(nrt) - 1
    want = np.zeros((T + 1, 30))
    for j in range(1, T + 1):
        rtl.ref_body_foot_traj(h, j, int(stop[j]), P(nrt[j].copy()), P(want[j]))
    st = np.zeros(138 + 6 * (nh + 2)); rtl.ref_body_foot_traj_state(h, P(st))
    rtl.ref_body_free(h)
    out, s_end, _ = oracle_rt_foot_run(oracle, nh, nrt, stop)
    np.testing.assert_array_equal(out[1:], want[1:, :6 * (nh + 1)])
    np.testing.assert_array_equal(s_end, st)


class _RtHooks(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("mod3", "foot", "rot", "body", "tx_total")]


class OracleRtNode:
    """The 100 Hz node on the oracle restatements: oracle/rt_glue.c with its orc_hook_* (rt_foot.c, foot_rot.c, ref_interp.c,
    body_mpc.c).  One instance = one robot."""

    def __init__(self, oracle, nh):
        lib = self.lib = oracle.lib
        self.nh = nh
        lib.orc_rt_node_doubles.restype = ctypes.c_int; lib.orc_rt_foot_state_doubles.restype = ctypes.c_int
        lib.orc_rt_ctx_bytes.restype = ctypes.c_int; lib.orc_body_mpc_bytes.restype = ctypes.c_int
        self.node = np.zeros(lib.orc_rt_node_doubles(nh)); lib.orc_rt_node_default(nh, P(self.node))
        self.foot = np.zeros(lib.orc_rt_foot_state_doubles(nh)); self.rot = np.zeros(2 + 6 * nh)
        self.body = ctypes.create_string_buffer(lib.orc_body_mpc_bytes())
        self.ctx = ctypes.create_string_buffer(lib.orc_rt_ctx_bytes())
        lib.orc_rt_ctx_init(self.ctx, nh, P(self.foot), P(self.rot), self.body)
        self.hk = _RtHooks(); lib.orc_rt_hooks_oracle(ctypes.byref(self.hk))

    def tick(self, msg, ctrl=1, bs=None):
        out = np.zeros(100)
        bs = np.zeros(4) if bs is None else bs
        self.lib.orc_rt_node_tick(self.nh, P(self.node), ctypes.byref(self.hk), self.ctx, P(np.ascontiguousarray(msg, dtype=float)), int(ctrl), P(bs), P(out))
        return out


def test_oracle_rt_node_lockstep_replay_bit_exact(oracle):
    """cfg1 lock-step replay: the 100 Hz node on the ORACLE restatements, fed with the /MPC/Gait messages the
    unmodified NLPRTControlClass::WalkingReactStepping published, against /rtMPC/traj of the unmodified PRMPCClass behind the
    same glue (tests/golden/rt_node_ref.npz, 1798 fast ticks, nh = 4): all 100 slots of every tick, bit for bit."""
    g = load("rt_node_ref.npz")
    nh = int(g["nh"][0])
    node = OracleRtNode(oracle, nh)
    msgs, want, mo = g["msgs"], g["out"], g["msg_of_fast"]
    for k in range(len(want)):
        got = node.tick(msgs[mo[k]])
        if not np.array_equal(got, want[k]):
            bad = np.nonzero(got != want[k])[0]
            raise AssertionError(f"fast tick {k}: slots {bad[:10]} differ: {got[bad[:4]]} vs {want[k][bad[:4]]}")
    np.testing.assert_array_equal(node.node, g["node_end"])


class OracleNlpNode:
    """The 40 Hz planner node on the oracle restatement (oracle/nlp_node.c)."""

    def __init__(self, oracle):
        lib = self.lib = oracle.lib
        lib.orc_nlp_node_doubles.restype = ctypes.c_int
        self.cfg = (ctypes.c_double * 128)()              # orc_nlp_cfg, opaque here (well under 1 KB)
        lib.orc_nlp_cfg_default(self.cfg)
        self.node = np.zeros(lib.orc_nlp_node_doubles())
        lib.orc_nlp_node_default(self.cfg, P(self.node))

    def start(self):
        self.lib.orc_nlp_node_start(P(self.node))

    def stop(self):
        self.lib.orc_nlp_node_stop(P(self.node))

    def tick(self, walkdtime, start_mpc, rf, lf):
        out = np.zeros(100)
        self.lib.orc_nlp_node_tick(self.cfg, P(self.node), int(walkdtime), int(start_mpc), P(np.ascontiguousarray(rf, dtype=float)),
                                   P(np.ascontiguousarray(lf, dtype=float)), P(out))
        return out


def run_nlp_script(node, T, events, idle, rf, lf):
    out = np.zeros((T + 1, 100))
    for count in range(1, T + 1):
        ev = events.get(count)
        if ev == "stop":
            node.stop()
        if ev == "start":
            node.start()
        out[count] = node.tick(count, 0 if count in idle else 1, rf[count], lf[count])
    return out


def test_oracle_nlp_node_bit_exact_vs_reference_golden(oracle):
    """NLPRTControlClass::WalkingReactStepping (+ StartWalking / StopWalking, X_CoM_position_squat, Zmp_distributor, the
    stop-walking branch of the swing foot) of the UNMODIFIED class: every /MPC/Gait slot of every tick, bit for bit, on cfg1's
    own message sequence (rt_node_ref.npz: 40 squat ticks, the 671-tick walk, 8 ticks beyond) and on scripted stop / restart /
    idle sequences with noisy foot-location feedback (nlp_node_ref.npz)."""
    from tests.golden.make_golden import nlp_node_feedback, nlp_node_scripts
    g = load("rt_node_ref.npz")
    msgs = g["msgs"]
    z = np.zeros((len(msgs), 3))
    got = run_nlp_script(OracleNlpNode(oracle), len(msgs) - 1, {}, (), z, z)
    np.testing.assert_array_equal(got[1:], msgs[1:])
    g2 = load("nlp_node_ref.npz")
    for name, T, events, idle, seed in nlp_node_scripts():
        rf, lf = nlp_node_feedback(seed, T)
        got = run_nlp_script(OracleNlpNode(oracle), T, events, idle, rf, lf)
        np.testing.assert_array_equal(got[1:], g2[name][1:], err_msg=name)
    assert g2["stop"][260:, [8, 11]].max() == 0.0 and g2["walk_fb"][260:, [8, 11]].max() > 0.03      # the stop zeroed the lift heights


def test_oracle_nlp_node_live_vs_reference(oracle):
    """Fresh stop / start scripts against the unmodified class where oracle/_ref is present (the authoring container)."""
    nlp = ref_path("libref_nlp.so")
    if not nlp:
        pytest.skip("oracle/_ref absent")
    nl = ctypes.CDLL(nlp)
    nl.ref_ctl_new.restype = ctypes.c_void_p
    rng = np.random.Generator(np.random.Philox(977))
    for trial in range(3):
        T = 500
        stop_at = int(rng.integers(60, 400)); start_at = stop_at + int(rng.integers(5, 80))
        events = {stop_at: "stop"} if trial == 0 else {stop_at: "stop", start_at: "start"}
        rf = np.zeros((T + 1, 3)); lf = np.zeros((T + 1, 3))
        rf[:, :2] = rng.uniform(-0.02, 0.02, (T + 1, 2)); lf[:, :2] = rng.uniform(-0.02, 0.02, (T + 1, 2))
        ctl = ctypes.c_void_p(nl.ref_ctl_new())
        node = OracleNlpNode(oracle)
        est = np.zeros(18)
        for count in range(1, T + 1):
            ev = events.get(count)
            if ev == "stop":
                nl.ref_ctl_stop(ctl); node.stop()
            if ev == "start":
                nl.ref_ctl_start(ctl); node.start()
            m = np.zeros(100)
            nl.ref_ctl_step(ctl, count, 1, P(est), P(rf[count].copy()), P(lf[count].copy()), P(m))
            m[98] = 0.0
            np.testing.assert_array_equal(node.tick(count, 1, rf[count], lf[count]), m, err_msg=f"trial {trial} tick {count}")
        nl.ref_ctl_free(ctl)
