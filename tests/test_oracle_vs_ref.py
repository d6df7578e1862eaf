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
