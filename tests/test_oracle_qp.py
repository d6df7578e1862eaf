"""Pins for the CPU oracle's QP solver (SURVEY.md section 4 list i-iv): known-answer
problems, KKT / dual-feasibility / complementarity checks, exact brute-force
active-set enumeration for n = 4, and a scipy cross-check."""
import itertools

import numpy as np
import pytest

from quadrupedal_loco_b200 import synth


def kkt_check(n, p, m, G, g0, CE, ce0, CI, ci0, x, active, tol=1e-7):
    """x minimises the QP iff it is feasible and -grad is a non-negative combination of the
    active inequality normals (+ any combination of equality normals)."""
    G = G.reshape(n, n, order="F"); CI = CI.reshape(n, m, order="F")
    CE = CE[:n * p].reshape(n, p, order="F") if p else np.zeros((n, 0))
    s = CI.T @ x + ci0
    assert s.min() > -tol, f"primal infeasible: {s.min()}"
    if p:
        assert np.abs(CE.T @ x + ce0[:p]).max() < tol
    grad = G @ x + g0
    ineq = [a for a in active if a >= 0]
    N = np.concatenate([CE, CI[:, ineq]], axis=1)
    if N.shape[1] == 0:
        assert np.abs(grad).max() < tol * max(1, np.abs(g0).max())
        return
    lam, *_ = np.linalg.lstsq(N, grad, rcond=None)
    assert np.abs(N @ lam - grad).max() < tol * max(1, np.abs(grad).max()), "stationarity"
    assert (lam[p:] > -tol).all(), "dual feasibility"
    assert np.abs(s[ineq]).max(initial=0) < tol, "complementarity"


def test_quadprogpp_demo(oracle):
    # the QuadProg++ lineage's demo problem (EiQuadProg.hpp:7-13 names the lineage)
    G = np.array([[4., -2], [-2, 4]]).ravel(order="F"); g0 = np.array([6., 0])
    CE = np.array([[1.], [1]]).ravel(order="F"); ce0 = np.array([-3.])
    CI = np.array([[1., 0, 1], [0, 1, 1]]).ravel(order="F"); ci0 = np.array([0., 0, -2])
    r = oracle.qp_solve(2, 1, 3, G, g0, CE, ce0, CI, ci0)
    assert r["status"] == 0
    np.testing.assert_allclose(r["x"], [1, 2], atol=1e-12)
    assert abs(r["cost"] - 12) < 1e-12
    assert list(r["active"]) == [-1]


def test_unconstrained_and_simple_bound(oracle):
    G = np.eye(3).ravel(order="F"); g0 = np.array([-1., -2, -3])
    CI = np.eye(3).ravel(order="F")
    r = oracle.qp_solve(3, 0, 3, G, g0, None, None, CI, np.array([0., 0, 0]))
    np.testing.assert_allclose(r["x"], [1, 2, 3], atol=1e-14)
    assert r["nactive"] == 0 and r["iters"][0] == 1
    # x >= 2.5 on the first coordinate binds
    r = oracle.qp_solve(3, 0, 3, G, g0, None, None, CI, np.array([-2.5, 0, 0]))
    np.testing.assert_allclose(r["x"], [2.5, 2, 3], atol=1e-14)
    assert list(r["active"]) == [0]


def test_not_pd_leaves_x_untouched(oracle):
    G = np.array([[1., 2], [2, 1]]).ravel(order="F")
    r = oracle.qp_solve(2, 0, 1, G, np.zeros(2), None, None, np.array([1., 0]), np.array([1.]), x0=[7., 8.])
    assert r["status"] == 1 and np.isinf(r["cost"])
    np.testing.assert_array_equal(r["x"], [7., 8.])


def test_infeasible(oracle):
    G = np.eye(2).ravel(order="F")
    CI = np.array([[1., 0], [-1., 0]]).T.ravel(order="F")   # x0 >= 1 and -x0 >= 1
    r = oracle.qp_solve(2, 0, 2, G, np.zeros(2), None, None, CI, np.array([-1., -1.]))
    assert r["status"] == 2 and np.isinf(r["cost"])


def test_zero_rows_never_activate(oracle):
    # never-populated constraint columns (body MPC rows 8nh..12nh) are zero rows with ci0 = 0
    G = np.eye(2).ravel(order="F"); g0 = np.array([1., 1.])
    CI = np.zeros((2, 3)); CI[:, 0] = [1, 0]
    r = oracle.qp_solve(2, 0, 3, G, g0, None, None, CI.ravel(order="F"), np.array([0., 0, 0]))
    np.testing.assert_allclose(r["x"], [0, -1], atol=1e-14)
    assert list(r["active"]) == [0]


@pytest.mark.parametrize("shape", [(4, 1, 24), (8, 0, 48), (20, 0, 120), (12, 2, 24)])
def test_random_kkt(oracle, shape):
    n, p, m = shape
    d = synth.random_qp(60, n, p, m, seed=11 + n, paired=True)
    for b in range(60):
        r = oracle.qp_solve(n, p, m, d["G"][b], d["g0"][b], d["CE"][b], d["ce0"][b], d["CI"][b], d["ci0"][b])
        assert r["status"] == 0
        kkt_check(n, p, m, d["G"][b], d["g0"][b], d["CE"][b], d["ce0"][b], d["CI"][b], d["ci0"][b], r["x"], r["active"])


def test_bruteforce_active_set_n4(oracle):
    """n = 4, m = 12: enumerate every active set of size <= 4 exactly (C(12,<=4) = 794 sets)."""
    n, m = 4, 12
    d = synth.random_qp(25, n, 0, m, seed=5)
    for b in range(25):
        G = d["G"][b].reshape(n, n, order="F"); g0 = d["g0"][b]
        CI = d["CI"][b].reshape(n, m, order="F"); ci0 = d["ci0"][b]
        best = (np.inf, None)
        for k in range(0, n + 1):
            for S in itertools.combinations(range(m), k):
                S = list(S)
                if k:
                    N = CI[:, S]
                    K = np.block([[G, -N], [N.T, np.zeros((k, k))]])
                    rhs = np.concatenate([-g0, -ci0[S]])
                    try:
                        sol = np.linalg.solve(K, rhs)
                    except np.linalg.LinAlgError:
                        continue
                    x, lam = sol[:n], sol[n:]
                    if (lam < -1e-9).any():
                        continue
                else:
                    x = np.linalg.solve(G, -g0)
                if (CI.T @ x + ci0 < -1e-9).any():
                    continue
                f = 0.5 * x @ G @ x + g0 @ x
                if f < best[0] - 1e-12:
                    best = (f, x)
        r = oracle.qp_solve(n, 0, m, d["G"][b], g0, None, None, d["CI"][b], ci0)
        assert r["status"] == 0
        np.testing.assert_allclose(r["x"], best[1], rtol=1e-8, atol=1e-9)
        assert abs(r["cost"] - best[0]) < 1e-8 * max(1, abs(best[0]))


def test_scipy_crosscheck(oracle):
    from scipy.optimize import minimize
    n, m = 6, 10
    d = synth.random_qp(8, n, 0, m, seed=3)
    for b in range(8):
        G = d["G"][b].reshape(n, n, order="F"); g0 = d["g0"][b]
        CI = d["CI"][b].reshape(n, m, order="F"); ci0 = d["ci0"][b]
        r = oracle.qp_solve(n, 0, m, d["G"][b], g0, None, None, d["CI"][b], ci0)
        res = minimize(lambda x: 0.5 * x @ G @ x + g0 @ x, np.zeros(n), jac=lambda x: G @ x + g0, method="SLSQP",
                       constraints=[{"type": "ineq", "fun": lambda x: CI.T @ x + ci0, "jac": lambda x: CI.T}],
                       options={"ftol": 1e-14, "maxiter": 500})
        np.testing.assert_allclose(r["x"], res.x, atol=2e-6)


def test_iteration_cap_reports_status(oracle):
    # sanity: a well-posed problem never reaches the cap, counters are consistent
    d = synth.random_qp(40, 8, 0, 48, seed=9, paired=True)
    for b in range(40):
        r = oracle.qp_solve(8, 0, 48, d["G"][b], d["g0"][b], None, None, d["CI"][b], d["ci0"][b])
        assert r["status"] == 0
        assert r["nactive"] == r["iters"][1] - r["iters"][2]


def test_native_thread_pool_matches_single_thread_drivers(oracle):
    """oracle/mt_pool.c (the CPU arm of bench.py): a 3-thread static split of a batch gives bit for bit what the
    single-threaded batch drivers give."""
    import ctypes
    from quadrupedal_loco_b200 import synth
    from tests.oracle_lib import P
    lib = oracle.lib
    B, nh = 101, 10
    d = synth.body_mpc_inputs(B, nh, seed=3)
    bcfg = oracle.body_cfg(nh); scfg = oracle.step_cfg(3)
    tick, st, sin = synth.step_timing_inputs(B, oracle.step_default_state(scfg), seed=3)
    theta = d["theta"].copy(); x = d["x_warm"].copy(); o14 = np.zeros((B, 14))
    oracle.body_step_batch(bcfg, d["tick"], d["tx"], theta, d["bstate"], d["refs"], o14, x)
    want38, _ = oracle.step_tick_batch(scfg, tick, st.copy(), sin)
    vp = ctypes.c_void_p
    lib.orc_pool_create.restype = vp
    lib.orc_pool_create.argtypes = [ctypes.c_int, vp, vp, ctypes.c_int] + [vp] * 8
    lib.orc_pool_run.restype = ctypes.c_double; lib.orc_pool_run.argtypes = [vp, ctypes.c_int]
    lib.orc_pool_results.argtypes = [vp, vp, vp]; lib.orc_pool_destroy.argtypes = [vp]
    keep = [np.ascontiguousarray(d["tick"], np.int32), np.ascontiguousarray(d["tx"]), np.ascontiguousarray(d["theta"]),
            np.ascontiguousarray(d["bstate"]), np.ascontiguousarray(d["refs"].reshape(B, 9 * nh)), np.ascontiguousarray(tick, np.int32),
            np.ascontiguousarray(st), np.ascontiguousarray(sin)]
    pool = lib.orc_pool_create(3, ctypes.addressof(bcfg), ctypes.addressof(scfg), B, *[P(k) for k in keep])
    assert pool
    assert lib.orc_pool_run(pool, 2) > 0
    g14 = np.zeros((B, 14)); g38 = np.zeros((B, 38))
    lib.orc_pool_results(pool, P(g14), P(g38))
    lib.orc_pool_destroy(pool)
    np.testing.assert_array_equal(g14, o14)
    np.testing.assert_array_equal(g38, want38)
