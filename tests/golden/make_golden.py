"""Generates tests/golden/*.npz from the UNMODIFIED reference sources compiled into
oracle/_ref/ (recipe: oracle/Makefile; needs /root/reference, so it only runs in the
authoring container).  The fixtures pin the oracle (tests/test_oracle_vs_ref.py) and give the
GPU tests reference-made outputs on the box, where /root/reference does not exist.

    python tests/golden/make_golden.py

Inputs are regenerated from seeds by quadrupedal_loco_b200.synth (stored too, so a change of
the generators cannot silently re-pin); outputs come from:
  ref_qp_solve        -> Eigen::QP::solve_quadprog        RT/src/utils/EiQuadProg/EiQuadProg.cpp:493-513
  ref_body_theta_mpc  -> PRMPCClass::body_theta_mpc       RT/src/FastMPC/PRMPCClass.cpp:379-714 (nh = 4)
  ref_fk/_g, ref_ik/_g-> Kinematicclass                   GO1/src/kinematics/Kinematics.cpp:63-304
  ref_nlp_step        -> NLPClass::step_timing_opti_loop  NLP/src/NLP/NLPClass_sqp.cpp:693-1102
"""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from quadrupedal_loco_b200 import synth  # noqa: E402
from tests.oracle_lib import P, PI, ref_path  # noqa: E402

QP_CASES = [  # (n, p, m, B, seed, kwargs)
    (4, 1, 24, 32, 1001, dict(paired=True)),
    (8, 0, 48, 32, 1002, dict(paired=True)),
    (12, 2, 24, 24, 1003, dict(paired=True)),
    (20, 0, 120, 16, 1004, dict(paired=True)),
    (6, 0, 14, 48, 1005, dict(dup=True, infeasible_frac=0.3)),
    (5, 0, 9, 24, 1006, dict()),
]


def gen_qp(lib):
    out = {}
    for ci, (n, p, m, B, seed, kw) in enumerate(QP_CASES):
        d = synth.random_qp(B, n, p, m, seed=seed, **kw)
        x = np.zeros((B, n)); cost = np.zeros(B); act = np.zeros((B, m + p + 1), np.int32); na = np.zeros(B, np.int32)
        for b in range(B):
            xb = np.zeros(n); cb = np.zeros(1); ab = np.zeros(m + p + 1, np.int32); nb = np.zeros(1, np.int32)
            lib.ref_qp_solve(n, p, m, P(d["G"][b].copy()), P(d["g0"][b].copy()), P(d["CE"][b].copy()), P(d["ce0"][b].copy()),
                             P(d["CI"][b].copy()), P(d["ci0"][b].copy()), P(xb), P(cb), PI(ab), PI(nb))
            x[b] = xb; cost[b] = cb[0]; act[b] = ab; na[b] = nb[0]
        for k in ("G", "g0", "CE", "ce0", "CI", "ci0"):
            out[f"c{ci}_{k}"] = d[k]
        out[f"c{ci}_x"] = x; out[f"c{ci}_cost"] = cost; out[f"c{ci}_active"] = act; out[f"c{ci}_nactive"] = na
        out[f"c{ci}_shape"] = np.array([n, p, m, B, seed])
    np.savez_compressed(os.path.join(HERE, "qp_ref.npz"), **out)
    print("qp_ref.npz", {k: v.shape for k, v in out.items() if k.endswith("_x")})


def gen_body(rt):
    nh = rt.ref_body_nh()
    assert nh == 4
    B, T = 48, 40
    d = synth.body_mpc_inputs(B, nh, seed=2001, scale=1.3)
    d["tick"][:] = 150 + 37 * np.arange(B)            # spread over the walk, crossing step switches
    d["tick"][:4] = [0, 50, 99, 100]                  # gated ticks and the first live one
    theta = d["theta"].copy(); vini = d["x_warm"].copy()
    o14 = np.zeros((T, B, 14)); th_t = np.zeros((T, B, 4)); v_t = np.zeros((T, B, 2 * nh)); ok = np.zeros((T, B), np.int32)
    tick0 = d["tick"].copy()
    rt.ref_body_new.restype = ctypes.c_void_p
    hs = [ctypes.c_void_p(rt.ref_body_new()) for _ in range(B)]
    for b in range(B):
        rt.ref_body_set_state(hs[b], P(d["tx"][b].copy()), P(theta[b].copy()), P(vini[b].copy()))
    for t in range(T):
        for b in range(B):
            r = d["refs"][b]
            o = np.zeros(14); q = np.zeros(1, np.int32)
            zmp = np.ascontiguousarray(r[0:2].ravel()); ang = np.ascontiguousarray(r[2:4].ravel())
            rf = np.ascontiguousarray(r[4:6].ravel()); lf = np.ascontiguousarray(r[6:8].ravel()); ca = np.ascontiguousarray(r[8])
            rt.ref_body_theta_mpc(hs[b], int(tick0[b] + t), P(d["bstate"][b].copy()), P(zmp), P(ang), P(rf), P(lf), P(ca), P(o), PI(q))
            tx = np.zeros(27); st = np.zeros(4); vi = np.zeros(2 * nh)
            rt.ref_body_get_state(hs[b], P(tx), P(st), P(vi))
            o14[t, b] = o; th_t[t, b] = st; v_t[t, b] = vi; ok[t, b] = q[0]
    for h in hs:
        rt.ref_body_free(h)
    np.savez_compressed(os.path.join(HERE, "body_ref_nh4.npz"), tick0=tick0, tx=d["tx"], theta0=d["theta"], bstate=d["bstate"],
                        refs=d["refs"], out14=o14, theta=th_t, vini=v_t, qp_ok=ok)
    print("body_ref_nh4.npz", o14.shape, "live ticks:", int((np.abs(o14).sum(axis=2) > 0).sum()))


def gen_kin(lib):
    rng = np.random.Generator(np.random.Philox(3001))
    N = 64
    q = np.stack([rng.uniform(-0.6, 0.6, N), rng.uniform(0.2, 1.4, N), rng.uniform(-2.2, -0.9, N)], 1)
    q[0] = [0, 0.6, -1.0]          # kinematics_matlab/forward_kin_go1.m demo pose
    q[1] = [0, 0.87, -1.5]         # homing pose, servo.cpp:767
    q[2] = [0, 0.67, -1.3]         # stand pose, body.cpp:42-43
    bp = rng.uniform(-0.05, 0.05, (N, 3)) + [0, 0, 0.31]; br = rng.uniform(-0.2, 0.2, (N, 3))
    leg = (np.arange(N) % 4).astype(np.int32)
    fk = np.zeros((N, 3)); fkJ = np.zeros((N, 9)); fkg = np.zeros((N, 3)); fkgJ = np.zeros((N, 9))
    for i in range(N):
        lib.ref_fk(P(q[i].copy()), int(leg[i]), P(fk[i]), P(fkJ[i]))
        lib.ref_fk_g(P(bp[i].copy()), P(br[i].copy()), P(q[i].copy()), int(leg[i]), P(fkg[i]), P(fkgJ[i]))
    # IK: targets = FK of the pose, start from a perturbed pose
    qini = q + rng.uniform(-0.15, 0.15, (N, 3))
    ik = np.zeros((N, 3)); ikJ = np.zeros((N, 9)); ikg = np.zeros((N, 3)); ikgJ = np.zeros((N, 9))
    for i in range(N):
        lib.ref_ik(P(fk[i].copy()), P(qini[i].copy()), int(leg[i]), P(ik[i]), P(ikJ[i]))
        lib.ref_ik_g(P(bp[i].copy()), P(br[i].copy()), P(fkg[i].copy()), P(qini[i].copy()), int(leg[i]), P(ikg[i]), P(ikgJ[i]))
    np.savez_compressed(os.path.join(HERE, "kin_ref.npz"), q=q, bp=bp, br=br, leg=leg, qini=qini, fk=fk, fkJ=fkJ, fkg=fkg,
                        fkgJ=fkgJ, ik=ik, ikJ=ikJ, ikg=ikg, ikgJ=ikgJ)
    print("kin_ref.npz", N)


def gen_step(nl):
    """NLPClass::step_timing_opti_loop: (a) the deterministic 671-tick replay (cfg1) sampled every
    tick; (b) one-tick evaluations from replay states with random pushes."""
    nl.ref_nlp_new.restype = ctypes.c_void_p
    nl.ref_nlp_new.argtypes = [ctypes.c_double] * 3
    S = 202
    est = np.zeros(18); rf = np.array([0, -0.12675, 0.]); lf = np.array([0, 0.12675, 0.])
    consts = np.zeros(30)

    def step(h, i):
        out = np.zeros(38); hz = np.zeros(10); ints = np.zeros(4, np.int32)
        nl.ref_nlp_step(h, i, P(est), P(rf), P(lf), 0, P(out), P(hz), PI(ints))
        inp = np.zeros(20); inp[6:8] = rf[:2]; inp[8:10] = lf[:2]; inp[10:13] = hz[0:3]; inp[13:16] = hz[3:6]
        inp[16:19] = hz[6:9]; inp[19] = hz[9]
        return out, inp, ints

    hA = ctypes.c_void_p(nl.ref_nlp_new(0.075, 0.2535, 0.0))
    nl.ref_nlp_consts(hA, P(consts))
    nl.ref_nlp_stepwidth0.restype = ctypes.c_double
    sw0 = nl.ref_nlp_stepwidth0(hA)
    T = 671
    st0 = np.zeros((T + 2, S)); outs = np.zeros((T + 1, 38)); ins = np.zeros((T + 1, 20)); ints = np.zeros((T + 1, 4), np.int32)
    foot = np.zeros((T + 1, 18)); foot_rs = np.zeros(T + 1, np.int32)
    for i in range(1, T + 1):
        nl.ref_nlp_get_state(hA, i, P(st0[i]))
        outs[i], ins[i], ints[i] = step(hA, i)
        foot_rs[i] = nl.ref_nlp_foot(hA, i, 0, P(foot[i]))     # NLPClass::Foot_trajectory_solve_mod2, same tick
    nl.ref_nlp_get_state(hA, T + 1, P(st0[T + 1]))
    # pushes
    hB = ctypes.c_void_p(nl.ref_nlp_new(0.075, 0.2535, 0.0))
    rng = np.random.Generator(np.random.Philox(4001))
    N = 384
    pt = rng.integers(2, 640, N).astype(np.int32)
    pst = st0[pt].copy()
    amp = np.where(np.arange(N) % 4 == 3, 1.2, 0.5)
    pst[:, 189] += amp * rng.uniform(-0.02, 0.02, N); pst[:, 190] += amp * rng.uniform(-0.25, 0.25, N)
    pst[:, 192] += amp * rng.uniform(-0.015, 0.015, N); pst[:, 193] += amp * rng.uniform(-0.2, 0.2, N)
    pout = np.zeros((N, 38)); pin = np.zeros((N, 20)); pints = np.zeros((N, 4), np.int32); pafter = np.zeros((N, S))
    for k in range(N):
        nl.ref_nlp_set_state(hB, int(pt[k]), P(pst[k].copy()))
        pout[k], pin[k], pints[k] = step(hB, int(pt[k]))
        nl.ref_nlp_get_state(hB, int(pt[k]) + 1, P(pafter[k]))
    np.savez_compressed(os.path.join(HERE, "step_ref.npz"), consts=consts, replay_state=st0, replay_out=outs, replay_in=ins,
                        replay_ints=ints, replay_foot=foot, replay_right_support=foot_rs, stepwidth0=np.array([sw0]), push_tick=pt, push_state=pst, push_in=pin, push_out=pout, push_ints=pints,
                        push_state_after=pafter)
    print("step_ref.npz replay", T, "pushes", N, "finite pushes", int(np.isfinite(pout).all(axis=1).sum()))


def gen_grf(dl):
    """Dynamiccclass::force_distribution + force_opt on 240 seeded cases (tests/test_grf.py: grf_inputs)."""
    from tests.test_grf import grf_inputs
    dl.ref_dyn_new.restype = ctypes.c_void_p
    h = ctypes.c_void_p(dl.ref_dyn_new())
    d = grf_inputs(240, seed=6); N = 240
    Fg = np.zeros((N, 12)); grf = d["prev"].copy(); ok = np.zeros(N, np.int32)
    for b in range(N):
        dl.ref_dyn_force_distribution(h, P(d["base"][b].copy()), P(d["legs"][b].copy()), P(d["F6"][b].copy()), int(d["mode"][b]),
                                      ctypes.c_double(0.9), P(d["rf"][b].copy()), P(d["lf"][b].copy()), P(Fg[b]))
        ok[b] = dl.ref_dyn_force_opt(h, P(d["base"][b].copy()), P(d["legs"][b].copy()), P(d["FT"][b].copy()), P(Fg[b]),
                                     int(d["mode"][b]), int(d["rs"][b]), ctypes.c_double(0.9), P(grf[b]))
    np.savez_compressed(os.path.join(HERE, "grf_ref.npz"), Fg=Fg, grf=grf, ok=ok, **d)
    print("grf_ref.npz", N)


def interp_inputs(N, seed=14):
    """Four consecutive 40 Hz samples of a 3-vector quantity per case and the 100 Hz tick they are interpolated at."""
    rng = np.random.Generator(np.random.Philox(seed))
    base = rng.uniform(-0.5, 0.5, (N, 1, 3)); vel = rng.uniform(-0.02, 0.02, (N, 1, 3))
    s = base + vel * np.arange(4).reshape(1, 4, 1) + rng.uniform(-0.002, 0.002, (N, 4, 3))
    wt = rng.integers(0, 2000, N).astype(np.int32)          # _t_end_footstep = 1610: some cases lie beyond it
    wt[::9] = rng.integers(0, 4, len(wt[::9]))          # the first ticks (t = 0: pow(0, 0))
    return dict(samples=s, walktime=wt)


def gen_interp(rtl):
    """PRMPCClass::XGetSolution_position_mod3 (+ _AAA_inv_mod) on 300 seeded cases."""
    rtl.ref_body_new.restype = ctypes.c_void_p
    h = ctypes.c_void_p(rtl.ref_body_new())
    d = interp_inputs(300); N = 300
    out = np.zeros((N, 21)); inv = np.zeros(16); tend = 0
    for b in range(N):
        s = d["samples"][b]
        tend = rtl.ref_body_position_mod3(h, int(d["walktime"][b]), ctypes.c_double(0.01), P(s[0].copy()), P(s[1].copy()), P(s[2].copy()),
                                          P(s[3].copy()), P(out[b]), P(inv))
    rtl.ref_body_free(h)
    np.savez_compressed(os.path.join(HERE, "interp_ref.npz"), out=out, inv=inv, t_end=np.array([tend]), nh=np.array([rtl.ref_body_nh()]), **d)
    print("interp_ref.npz", N, "t_end_footstep", tend, "beyond", int((d["walktime"] > tend).sum()))


def rt_foot_inputs(seed=16):
    """100 Hz call sequence of PRMPCClass::Foot_trajectory_solve_mod2: tick j = 1..1750 with the Nrtfoorpr_gen (message slots
    86-94 = Vec38[27:36]) of the 40 Hz planner replay (step_ref.npz) that is current at that time, the planner's step
    periods and landing positions jittered so that the swing cubics, the landing window and re-timed steps are exercised;
    stop-walking is raised for the last 60 ticks."""
    g = np.load(os.path.join(HERE, "step_ref.npz"))
    rng = np.random.Generator(np.random.Philox(seed))
    T = 1750
    out38 = g["replay_out"]
    nrt = np.zeros((T + 1, 9)); stop = np.zeros(T + 1, np.int32)
    jit = rng.uniform(-0.01, 0.01, (len(out38), 9))
    jit[:, 0] = 0; jit[:, 7] = 0; jit[:, 8] = rng.uniform(-0.05, 0.05, len(out38))
    for j in range(1, T + 1):
        m = min(len(out38) - 1, 1 + int(np.floor(j * 0.01 / 0.025)))
        nrt[j] = out38[m, 27:36] + jit[m]
    stop[T - 60:] = 1
    return nrt, stop


def gen_rt_foot(rtl):
    """PRMPCClass::Foot_trajectory_solve_mod2 on one object over the 1750-call sequence (tests/test_oracle_vs_ref.py)."""
    rtl.ref_body_new.restype = ctypes.c_void_p
    h = ctypes.c_void_p(rtl.ref_body_new())
    nh = rtl.ref_body_nh()
    nrt, stop = rt_foot_inputs()
    T = len(nrt) - 1
    out = np.zeros((T + 1, 30)); S = 138 + 6 * (nh + 2)
    st = np.zeros((T + 1, S))
    rtl.ref_body_foot_traj_state(h, P(st[0]))
    for j in range(1, T + 1):
        rtl.ref_body_foot_traj(h, j, int(stop[j]), P(nrt[j].copy()), P(out[j]))
        rtl.ref_body_foot_traj_state(h, P(st[j]))
    rtl.ref_body_free(h)
    np.savez_compressed(os.path.join(HERE, "rt_foot_ref.npz"), nrt=nrt, stop=stop, out=out, state0=st[0], state_end=st[T],
                        state_mid=st[T // 2], nh=np.array([nh]))
    print("rt_foot_ref.npz", T, "moving samples", int((np.abs(np.diff(out[:, 0])) > 0).sum()))


class RtHooks(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("mod3", "foot", "rot", "body", "tx_total")]


def ref_rt_hooks(rtl):
    hk = RtHooks()
    for n in ("mod3", "foot", "rot", "body", "tx_total"):
        setattr(hk, n, ctypes.cast(getattr(rtl, "ref_hook_" + n), ctypes.c_void_p).value)
    return hk


def gen_rt_node(rtl, nl, orc):
    """cfg1 lock-step replay of the two MPC nodes from the UNMODIFIED classes: NLPRTControlClass::WalkingReactStepping at 40 Hz
    (40 squat ticks + the 671-tick walk) publishes the 100-slot /MPC/Gait message; the 100 Hz node -- gait_fast.cpp's glue
    (restated once in oracle/rt_glue.c) around the unmodified PRMPCClass -- consumes the latest message every 10 ms and
    publishes /rtMPC/traj.  Slow ticks at t = 25 k ms run before the fast tick of the same instant."""
    nl.ref_ctl_new.restype = ctypes.c_void_p
    rtl.ref_body_new.restype = ctypes.c_void_p
    ctl = ctypes.c_void_p(nl.ref_ctl_new()); body = ctypes.c_void_p(rtl.ref_body_new())
    nh = rtl.ref_body_nh()
    orc.orc_rt_node_doubles.restype = ctypes.c_int
    node = np.zeros(orc.orc_rt_node_doubles(nh)); orc.orc_rt_node_default(nh, P(node))
    hk = ref_rt_hooks(rtl)
    est = np.zeros(18); rf = np.zeros(3); lf = np.zeros(3); bs = np.zeros(4)
    n_slow = 40 + 671 + 8
    msgs = np.zeros((n_slow + 1, 100)); outs = []; msg_of_fast = []
    msg = np.zeros(100); count = 0
    t_ms = 0
    while count < n_slow or t_ms % 25:
        if t_ms % 25 == 0:
            count += 1
            m = np.zeros(100)
            nl.ref_ctl_step(ctl, count, 1, P(est), P(rf), P(lf), P(m))
            m[98] = 0.0                                        # wall time of the tick
            msgs[count] = m; msg = m
        if t_ms % 10 == 0:
            o = np.zeros(100)
            orc.orc_rt_node_tick(nh, P(node), ctypes.byref(hk), body, P(msg.copy()), 1, P(bs), P(o))
            outs.append(o); msg_of_fast.append(count)
        t_ms += 5
    nl.ref_ctl_free(ctl); rtl.ref_body_free(body)
    outs = np.array(outs)
    np.savez_compressed(os.path.join(HERE, "rt_node_ref.npz"), msgs=msgs, out=outs, msg_of_fast=np.array(msg_of_fast, np.int32),
                        node_end=node, nh=np.array([nh]))
    live = np.abs(outs[:, 72:86]).sum(axis=1) > 0
    print("rt_node_ref.npz slow", count, "fast", len(outs), "fast ticks with a body-MPC result", int(live.sum()))


def foot_rot_inputs(seed=15):
    """Step-length table with forward, backward and zero-length steps, and a tick sequence that walks the whole table,
    runs past _t_end_footstep and jumps back (the angle members persist between calls)."""
    rng = np.random.Generator(np.random.Philox(seed))
    steps = rng.uniform(0.02, 0.12, 27); steps[[6, 15]] = 0.0; steps[9:12] *= -1
    footx = np.concatenate([[0.0], np.cumsum(steps)[:-1]])
    ticks = np.concatenate([np.arange(1, 1700, 3), [1650, 1700, 20, 500, 499, 1611, 1610]]).astype(np.int32)
    return footx, ticks


def gen_foot_rot(rtl):
    """PRMPCClass::XGetSolution_Foot_rotation over one object, sequential calls (tests/test_oracle_vs_ref.py)."""
    rtl.ref_body_new.restype = ctypes.c_void_p
    h = ctypes.c_void_p(rtl.ref_body_new())
    tx = np.zeros(27); ts = np.zeros(27); td = np.zeros(27); fx0 = np.zeros(27); sc = np.zeros(4)
    rtl.ref_body_foot_tables(h, P(tx), P(ts), P(td), P(fx0), P(sc))
    footx, ticks = foot_rot_inputs()
    rtl.ref_body_set_footx(h, P(footx.copy()))
    out = np.zeros((len(ticks), 30))
    for k, t in enumerate(ticks):
        rtl.ref_body_foot_rotation(h, int(t), ctypes.c_double(0.01), P(out[k]))
    nh = rtl.ref_body_nh()
    rtl.ref_body_free(h)
    np.savez_compressed(os.path.join(HERE, "foot_rot_ref.npz"), tx=tx, ts=ts, td=td, footx=footx, scal=sc, ticks=ticks, out=out, nh=np.array([nh]))
    print("foot_rot_ref.npz", len(ticks), "non-zero rows", int((np.abs(out).sum(axis=1) > 0).sum()))


def gen_grf_tau(dl):
    """Dynamiccclass::compute_joint_torques on 400 seeded legs (tests/test_grf.py: tau_inputs)."""
    from tests.test_grf import tau_inputs
    dl.ref_dyn_new.restype = ctypes.c_void_p
    h = ctypes.c_void_p(dl.ref_dyn_new())
    d = tau_inputs(100, seed=8); N = 100
    tau = np.zeros((N, 4, 3))
    for b in range(N):
        for leg in range(4):
            dl.ref_dyn_joint_torques(h, P(d["jac"][b, leg].copy()), int(d["swing"][b, leg]), P(d["p_des"][b, leg].copy()), P(d["p_est"][b, leg].copy()),
                                     P(d["pv_des"][b, leg].copy()), P(d["pv_est"][b, leg].copy()), P(d["F"][b, leg].copy()), leg, P(tau[b, leg]))
    np.savez_compressed(os.path.join(HERE, "grf_tau_ref.npz"), tau=tau, **d)
    print("grf_tau_ref.npz", N)


def filter_inputs(seed=41, T=400, C=6):
    rng = np.random.Generator(np.random.Philox(seed))
    t = np.arange(T)[:, None]
    return np.sin(0.013 * t * (1 + np.arange(C))) + 0.2 * rng.standard_normal((T, C)) + np.arange(C)


def gen_filters(rf):
    """butterworthLPF (the servo's (fs, fc) pairs: 1 kHz with 3 / 10 / 20 Hz cut-offs) and ButterworthFilter::ForceFilter of the
    UNMODIFIED classes on seeded noisy signals."""
    rf.ref_lpf_new.restype = ctypes.c_void_p; rf.ref_lpf_new.argtypes = [ctypes.c_double] * 2
    rf.ref_lpf_filter.restype = ctypes.c_double; rf.ref_lpf_filter.argtypes = [ctypes.c_void_p, ctypes.c_double]
    rf.ref_lpf_coefs.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    rf.ref_force_filter_new.restype = ctypes.c_void_p
    rf.ref_force_filter.restype = ctypes.c_double; rf.ref_force_filter.argtypes = [ctypes.c_void_p, ctypes.c_double]
    x = filter_inputs()
    T, C = x.shape
    fsfc = np.array([(1000.0, 3.0), (1000.0, 10.0), (1000.0, 20.0), (400.0, 5.0), (1000.0, 3.0), (200.0, 30.0)])
    y = np.zeros((T, C)); coefs = np.zeros((C, 7)); yf = np.zeros((T, C))
    for c in range(C):
        h = ctypes.c_void_p(rf.ref_lpf_new(*fsfc[c])); rf.ref_lpf_coefs(h, P(coefs[c]))
        g = ctypes.c_void_p(rf.ref_force_filter_new())
        for t in range(T):
            y[t, c] = rf.ref_lpf_filter(h, float(x[t, c])); yf[t, c] = rf.ref_force_filter(g, float(50 * x[t, c]))
    np.savez_compressed(os.path.join(HERE, "filter_ref.npz"), x=x, fsfc=fsfc, coefs=coefs, lpf=y, force=yf)
    print("filter_ref.npz", T, C)


def nlp_node_scripts():
    """Scripted sequences for the 40 Hz node: (name, ticks, {tick: 'stop' | 'start'}, ticks with start_mpc = 0, foot-feedback seed)."""
    return [("walk_fb", 739, {}, (), 21),                     # cfg1 with noisy foot-location feedback, runs past the end
            ("stop_early", 120, {45: "stop"}, (), 0),         # StopWalking while _t_int < 10: no effect on the flags
            ("stop", 420, {200: "stop"}, (), 22),             # lift heights of the steps ahead zeroed
            ("stop_restart", 420, {200: "stop", 300: "start"}, (), 23),   # _start_walking_again: the node freezes
            ("idle", 200, {}, tuple(range(1, 25)) + (30, 31), 24)]       # start_mpc = 0 before / during the squat (a gap in the WALK
                                                                          # ticks is outside the contract: the reference then
                                                                          # reads whole-walk array entries no tick ever wrote)


def nlp_node_feedback(seed, T):
    rng = np.random.Generator(np.random.Philox(5000 + seed))
    rf = np.zeros((T + 1, 3)); lf = np.zeros((T + 1, 3))
    if seed:
        rf[:, :2] = rng.uniform(-0.01, 0.01, (T + 1, 2)); lf[:, :2] = rng.uniform(-0.01, 0.01, (T + 1, 2))
    return rf, lf


def gen_nlp_node(nl):
    """NLPRTControlClass::WalkingReactStepping / StartWalking / StopWalking of the UNMODIFIED class on scripted sequences."""
    nl.ref_ctl_new.restype = ctypes.c_void_p
    out = {}
    for name, T, events, idle, seed in nlp_node_scripts():
        ctl = ctypes.c_void_p(nl.ref_ctl_new())
        rf, lf = nlp_node_feedback(seed, T)
        msgs = np.zeros((T + 1, 100)); est = np.zeros(18)
        for count in range(1, T + 1):
            ev = events.get(count)
            if ev == "stop": nl.ref_ctl_stop(ctl)
            if ev == "start": nl.ref_ctl_start(ctl)
            nl.ref_ctl_step(ctl, count, 0 if count in idle else 1, P(est), P(rf[count].copy()), P(lf[count].copy()), P(msgs[count]))
            msgs[count, 98] = 0.0
        nl.ref_ctl_free(ctl)
        out[name] = msgs
        print("nlp_node_ref.npz", name, T, "NaN slots", int(np.isnan(msgs).sum()))
    np.savez_compressed(os.path.join(HERE, "nlp_node_ref.npz"), **out)


if __name__ == "__main__":
    ref, rt, nlp = ref_path("libref.so"), ref_path("libref_rt.so"), ref_path("libref_nlp.so")
    if not ref or not rt or not nlp:
        raise SystemExit("oracle/_ref is missing: run `make -C oracle ref` where /root/reference exists")
    lib = ctypes.CDLL(ref); rtl = ctypes.CDLL(rt)
    only = sys.argv[1:]
    if not only or "qp" in only:
        gen_qp(lib)
    if not only or "body" in only:
        gen_body(rtl)
    if not only or "kin" in only:
        gen_kin(lib)
    if not only or "step" in only:
        gen_step(ctypes.CDLL(nlp))
    if not only or "grf" in only:
        gen_grf(ctypes.CDLL(ref_path("libref_dyn.so")))
    if not only or "foot_rot" in only:
        gen_foot_rot(rtl)
    if not only or "interp" in only:
        gen_interp(rtl)
    if not only or "rt_foot" in only:
        gen_rt_foot(rtl)
    if not only or "rt_node" in only:
        gen_rt_node(rtl, ctypes.CDLL(nlp), ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle.so")))
    if not only or "filters" in only:
        gen_filters(ctypes.CDLL(ref_path("libref_filter.so")))
    if not only or "nlp_node" in only:
        gen_nlp_node(ctypes.CDLL(nlp))
    if not only or "grf_tau" in only:
        gen_grf_tau(ctypes.CDLL(ref_path("libref_dyn.so")))
