"""Ground-reaction-force distribution of go1_servo: the oracle against
the unmodified Dynamiccclass (CPU, live where oracle/_ref exists + golden vectors) and the CUDA
kernels against the oracle and the golden vectors (GPU).  Reference:
GO1/src/whole_body_dynamics/dynmics_compute.cpp:141-427."""
import ctypes
import os

import numpy as np
import pytest

from tests.oracle_lib import P, PI, ref_path
from tests.test_oracle_vs_ref import load

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "grf_ref.npz")


class GrfCfg(ctypes.Structure):
    _fields_ = [(n, ctypes.c_double) for n in "qp_alpha qp_beta qp_gama fz_max mu".split()]


def grf_inputs(N, seed=6):
    rng = np.random.Generator(np.random.Philox(seed))
    hom = np.array([0.1881, -0.1268, 0, 0.1881, 0.1268, 0, -0.1881, -0.1268, 0, -0.1881, 0.1268, 0])
    d = dict(mode=(101 + np.arange(N) % 2).astype(np.int32), rs=(np.arange(N) % 3).astype(np.int32),
             base=np.array([0, 0, 0.3]) + rng.uniform(-0.03, 0.03, (N, 3)), legs=hom + rng.uniform(-0.03, 0.03, (N, 12)))
    d["FT"] = np.stack([rng.uniform(-15, 15, N), rng.uniform(-15, 15, N), 117.6 + rng.uniform(-20, 20, N), rng.uniform(-3, 3, N),
                        rng.uniform(-3, 3, N), rng.uniform(-2, 2, N)], 1)
    d["F6"] = np.stack([rng.uniform(-8, 8, N), rng.uniform(-8, 8, N), rng.uniform(20, 100, N), rng.uniform(-8, 8, N),
                        rng.uniform(-8, 8, N), rng.uniform(20, 100, N)], 1)
    d["rf"] = np.array([0.0, -0.127, 0]) + rng.uniform(-0.03, 0.03, (N, 3)); d["lf"] = np.array([0.0, 0.127, 0]) + rng.uniform(-0.03, 0.03, (N, 3))
    d["prev"] = rng.uniform(-5, 40, (N, 12))
    return d


def oracle_grf(oracle, d, yc=0.9):
    N = len(d["mode"])
    cfg = GrfCfg(); oracle.lib.orc_grf_cfg_default(ctypes.byref(cfg))
    oracle.lib.orc_grf_force_distribution.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int, ctypes.c_double] + [ctypes.c_void_p] * 3
    Fg = np.zeros((N, 12)); grf = d["prev"].copy(); st = np.zeros(N, np.int32); na = np.zeros(N, np.int32)
    act = np.zeros((N, 40), np.int32); it = np.zeros((N, 4), np.int32); ok = np.zeros(N, np.int32)
    for b in range(N):
        oracle.lib.orc_grf_force_distribution(P(d["base"][b].copy()), P(d["legs"][b].copy()), P(d["F6"][b].copy()), int(d["mode"][b]), yc,
                                              P(d["rf"][b].copy()), P(d["lf"][b].copy()), P(Fg[b]))
        st[b] = oracle.lib.orc_grf_force_opt(ctypes.byref(cfg), P(d["base"][b].copy()), P(d["legs"][b].copy()), P(d["FT"][b].copy()), P(Fg[b]),
                                             int(d["mode"][b]), int(d["rs"][b]), P(grf[b]), PI(act[b]), PI(na[b:b + 1]), PI(it[b]), PI(ok[b:b + 1]))
    return dict(Fg=Fg, grf=grf, status=st, nactive=na, active=act, iters=it, ok=ok)


def test_oracle_grf_vs_reference_golden(oracle):
    g = np.load(GOLD)
    d = {k: g[k] for k in ("mode", "rs", "base", "legs", "FT", "F6", "rf", "lf", "prev")}
    o = oracle_grf(oracle, d)
    np.testing.assert_array_equal(o["Fg"], g["Fg"])
    np.testing.assert_array_equal(o["grf"], g["grf"])
    np.testing.assert_array_equal(o["ok"], g["ok"])
    assert (o["nactive"] >= 6).any() and (o["nactive"] == 0).any()     # swing-leg equalities present / absent


@pytest.mark.skipif(ref_path("libref_dyn.so") is None, reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_grf_vs_live_reference(oracle):
    lib = ctypes.CDLL(ref_path("libref_dyn.so")); lib.ref_dyn_new.restype = ctypes.c_void_p
    h = ctypes.c_void_p(lib.ref_dyn_new())
    d = grf_inputs(200, seed=77)
    o = oracle_grf(oracle, d)
    for b in range(200):
        Fr = np.zeros(12)
        lib.ref_dyn_force_distribution(h, P(d["base"][b].copy()), P(d["legs"][b].copy()), P(d["F6"][b].copy()), int(d["mode"][b]),
                                       ctypes.c_double(0.9), P(d["rf"][b].copy()), P(d["lf"][b].copy()), P(Fr))
        g1 = d["prev"][b].copy()
        ok = lib.ref_dyn_force_opt(h, P(d["base"][b].copy()), P(d["legs"][b].copy()), P(d["FT"][b].copy()), P(Fr), int(d["mode"][b]),
                                   int(d["rs"][b]), ctypes.c_double(0.9), P(g1))
        np.testing.assert_array_equal(Fr, o["Fg"][b]); np.testing.assert_array_equal(g1, o["grf"][b]); assert ok == o["ok"][b]
    lib.ref_dyn_free(h)


GOLD_TAU = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "grf_tau_ref.npz")
GRAV = np.array([[-0.80, 0, 0], [0.80, 0, 0], [-0.80, 0, 0], [0.80, 0, 0]])     # gravity_compensate columns, dynmics_compute.cpp:39-41


def tau_inputs(N, seed=8):
    """Per robot and leg: a leg Jacobian (row-major 3x3), swing flag, desired / estimated foot position and velocity, leg force."""
    rng = np.random.Generator(np.random.Philox(seed))
    d = dict(jac=rng.uniform(-0.3, 0.3, (N, 4, 9)), swing=(rng.uniform(size=(N, 4)) < 0.4).astype(np.int32),
             p_des=rng.uniform(-0.3, 0.3, (N, 4, 3)), pv_des=rng.uniform(-1, 1, (N, 4, 3)), F=rng.uniform(-20, 80, (N, 4, 3)))
    d["p_est"] = d["p_des"] + rng.uniform(-0.02, 0.02, (N, 4, 3)); d["pv_est"] = d["pv_des"] + rng.uniform(-0.3, 0.3, (N, 4, 3))
    d["jac"][::7, :, 4] = 0.0          # exact zeros in the Jacobian, as hip columns have
    return d


def oracle_tau(oracle, d):
    N = len(d["swing"])
    f = oracle.lib.orc_grf_joint_torques
    f.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 6 + [ctypes.c_double, ctypes.c_double, ctypes.c_void_p]
    f.restype = None
    tau = np.zeros((N, 4, 3))
    for b in range(N):
        for leg in range(4):
            f(P(d["jac"][b, leg].copy()), int(d["swing"][b, leg]), P(d["p_des"][b, leg].copy()), P(d["p_est"][b, leg].copy()),
              P(d["pv_des"][b, leg].copy()), P(d["pv_est"][b, leg].copy()), P(d["F"][b, leg].copy()), P(GRAV[leg].copy()), 1.0, 0.01, P(tau[b, leg]))
    return tau


def test_oracle_joint_torques_vs_reference_golden(oracle):
    """orc_grf_joint_torques against Dynamiccclass::compute_joint_torques (golden vectors of the unmodified class): bit-exact."""
    g = np.load(GOLD_TAU)
    d = {k: g[k] for k in ("jac", "swing", "p_des", "p_est", "pv_des", "pv_est", "F")}
    np.testing.assert_array_equal(oracle_tau(oracle, d), g["tau"])
    assert d["swing"].any() and not d["swing"].all()


@pytest.mark.skipif(ref_path("libref_dyn.so") is None, reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_joint_torques_vs_live_reference(oracle):
    lib = ctypes.CDLL(ref_path("libref_dyn.so")); lib.ref_dyn_new.restype = ctypes.c_void_p
    if not hasattr(lib, "ref_dyn_joint_torques"):
        pytest.skip("oracle/_ref predates ref_dyn_joint_torques")
    h = ctypes.c_void_p(lib.ref_dyn_new())
    d = tau_inputs(150, seed=81)
    want = oracle_tau(oracle, d)
    t = np.zeros(3)
    for b in range(150):
        for leg in range(4):
            lib.ref_dyn_joint_torques(h, P(d["jac"][b, leg].copy()), int(d["swing"][b, leg]), P(d["p_des"][b, leg].copy()), P(d["p_est"][b, leg].copy()),
                                      P(d["pv_des"][b, leg].copy()), P(d["pv_est"][b, leg].copy()), P(d["F"][b, leg].copy()), leg, P(t))
            np.testing.assert_array_equal(t, want[b, leg])
    lib.ref_dyn_free(h)


@pytest.mark.gpu
def test_gpu_joint_torques_vs_oracle_and_golden(mpc, oracle):
    """go1mpc_grf_joint_torques_batch: bit-exact against the oracle and the reference's golden vectors, with F_leg_ref read
    from the SoA layout and from force_opt-style out records."""
    import torch
    dev = torch.device("cuda", 0)
    g = np.load(GOLD_TAU)
    for src in ("golden", "synth"):
        d = {k: g[k] for k in ("jac", "swing", "p_des", "p_est", "pv_des", "pv_est", "F")} if src == "golden" else tau_inputs(5000, seed=12)
        N = len(d["swing"])
        want = g["tau"] if src == "golden" else oracle_tau(oracle, d)
        soa = lambda a: torch.from_numpy(np.array(a.reshape(N, -1).T, order="C", copy=True)).to(dev)
        args = [soa(d[k]) for k in ("jac", "swing", "p_des", "p_est", "pv_des", "pv_est")]
        F_soa = soa(d["F"])
        F_rec = torch.zeros(N, 16, dtype=torch.float64, device=dev); F_rec[:, :12] = torch.from_numpy(d["F"].reshape(N, 12)).to(dev)
        for F, strides in ((F_soa, None), (F_rec, (1, 16))):
            tau = torch.full((12, N), np.nan, dtype=torch.float64, device=dev)
            torch.cuda.synchronize()
            mpc.grf_joint_torques(N, *args, F, tau, F_strides=strides)
            mpc.synchronize()
            np.testing.assert_array_equal(tau.cpu().numpy().T.reshape(N, 4, 3), want, err_msg=f"{src} {strides}")


@pytest.mark.gpu
def test_gpu_grf_vs_oracle_and_golden(mpc, oracle):
    import torch
    dev = torch.device("cuda", 0)
    g = np.load(GOLD)
    for src in ("golden", "synth"):
        d = {k: g[k] for k in ("mode", "rs", "base", "legs", "FT", "F6", "rf", "lf", "prev")} if src == "golden" else grf_inputs(3000, seed=9)
        N = len(d["mode"])
        o = oracle_grf(oracle, d)
        # closed-form split on the device (one launch per gait mode, as the ABI takes a single mode)
        Fg = np.zeros((N, 12))
        soa = lambda a: torch.from_numpy(np.array(a.T, order="C", copy=True)).to(dev)
        for mode in (101, 102):
            sel = np.nonzero(d["mode"] == mode)[0]
            out = torch.zeros(12, len(sel), dtype=torch.float64, device=dev)
            torch.cuda.synchronize()
            mpc.grf_force_distribution(len(sel), mode, 0.9, soa(d["base"][sel]), soa(d["legs"][sel]), soa(d["F6"][sel]), soa(d["rf"][sel]),
                                       soa(d["lf"][sel]), out)
            mpc.synchronize()
            Fg[sel] = out.cpu().numpy().T
        np.testing.assert_allclose(Fg, o["Fg"], rtol=1e-12, atol=1e-12)
        rec = np.zeros((N, 48))
        rec[:, 0:3] = d["base"]; rec[:, 3:15] = d["legs"]; rec[:, 15:21] = d["FT"]; rec[:, 21:33] = o["Fg"]; rec[:, 33:45] = d["prev"]
        rec[:, 45] = d["mode"]; rec[:, 46] = d["rs"]
        tin = torch.from_numpy(rec).to(dev); tout = torch.zeros(N, 16, dtype=torch.float64, device=dev)
        tdg = torch.zeros(N, 32, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        mpc.grf_force_opt(N, tin, tout, tdg)
        mpc.synchronize()
        out = tout.cpu().numpy(); dg = tdg.cpu().numpy()
        assert np.array_equal(dg[:, 0], o["status"]), src
        conv = o["status"] == 0
        assert np.array_equal(dg[conv, 1], o["nactive"][conv]) and np.array_equal(dg[conv, 2:6], o["iters"][conv])
        for b in np.nonzero(conv)[0]:
            k = o["nactive"][b]
            assert np.array_equal(dg[b, 8:8 + k], o["active"][b, :k]), (src, b)
        sc = np.maximum(1.0, np.abs(o["grf"]).max(axis=1, keepdims=True))
        assert (np.abs(out[:, :12] - o["grf"]) / sc)[conv].max() < 1e-9, src
        if src == "golden":
            assert (np.abs(out[:, :12] - g["grf"]) / sc)[conv].max() < 1e-9
        # physics of converged solves: unilateral contact, friction pyramid, swing legs force-free
        f = out[conv, :12].reshape(-1, 4, 3)
        assert (f[:, :, 2] > -1e-7).all() and (f[:, :, 2] < 160 + 1e-7).all()
        assert (np.abs(f[:, :, 0]) <= 0.25 * f[:, :, 2] + 1e-7).all() and (np.abs(f[:, :, 1]) <= 0.25 * f[:, :, 2] + 1e-7).all()


@pytest.mark.gpu
def test_gpu_joint_torques_edge_cases(mpc, oracle):
    """Empty batch = no-op, missing buffers = error, one robot through the host form = the oracle."""
    import quadrupedal_loco_b200 as q
    assert mpc.lib.go1mpc_grf_joint_torques_batch_host(mpc.h, 0, None, None, None, None, None, None, None, 1, 1, None) == 0
    d = tau_inputs(1, seed=3)
    want = oracle_tau(oracle, d)
    flat = lambda a: np.ascontiguousarray(a.reshape(1, -1).T)
    tau = np.full((12, 1), np.nan)
    with pytest.raises(q.Go1MpcError):
        mpc._check(mpc.lib.go1mpc_grf_joint_torques_batch_host(mpc.h, 1, None, None, None, None, None, None, None, 1, 1, tau.ctypes.data), "null")
    args = [flat(d[k]) for k in ("jac", "swing", "p_des", "p_est", "pv_des", "pv_est", "F")]
    rc = mpc.lib.go1mpc_grf_joint_torques_batch_host(mpc.h, 1, *[a.ctypes.data for a in args], 1, 1, tau.ctypes.data)
    assert rc == 0
    np.testing.assert_array_equal(tau.T.reshape(1, 4, 3), want)
