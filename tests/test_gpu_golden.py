"""GPU path against outputs produced by the UNMODIFIED reference (tests/golden/*.npz, made by
tests/golden/make_golden.py in the authoring container): primal 1e-9 relative, identical
final active set for converged solves."""
import os

import numpy as np
import pytest

import quadrupedal_loco_b200 as q
from tests.test_gpu_qp import gpu_solve
from tests.test_oracle_vs_ref import load, qp_cases

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def test_gpu_qp_vs_reference_golden(mpc):
    for ci, n, p, m, B, d in qp_cases():
        g = gpu_solve(mpc, n, p, m, d)
        ok = ~np.isinf(d["cost"])
        assert np.array_equal(np.isinf(g["cost"]), ~ok), ci
        scale = np.maximum(1.0, np.abs(d["x"]).max(axis=1, keepdims=True))
        err = (np.abs(g["x"] - d["x"]) / scale)[ok]
        assert err.max() < RTOL, (ci, err.max())
        for b in np.nonzero(ok)[0]:
            k = d["nactive"][b]
            assert g["nactive"][b] == k and np.array_equal(g["active"][b, :k], d["active"][b, :k]), (ci, b)
        assert (np.abs(g["cost"][ok] - d["cost"][ok]) / np.maximum(1, np.abs(d["cost"][ok]))).max() < 1e-8


def test_gpu_body_vs_reference_golden_nh4(mpc):
    """Closed loop at the reference's own horizon (nh = 4): 40 ticks x 48 instances, GPU state
    carried on the GPU side, compared with PRMPCClass::body_theta_mpc's outputs every tick."""
    g = load("body_ref_nh4.npz")
    nh = 4
    T, B = g["out14"].shape[:2]
    theta = g["theta0"].copy(); x = np.zeros((B, 2 * nh)); o14 = np.zeros((B, 14))
    for t in range(T):
        rec = q.pack_body_inputs(nh, g["tick0"] + t, g["tx"], theta, g["bstate"], x, g["refs"])
        out = np.zeros((B, q.body_out_stride(nh))); out[:, :14] = o14
        diag = np.zeros((B, q.body_diag_stride(nh)), np.int32)
        mpc.body_mpc_step_host(nh, B, rec, out, diag)
        o14 = out[:, :14].copy(); theta = out[:, 14:18].copy(); x = out[:, 18:18 + 2 * nh].copy()
        sc = np.maximum(1.0, np.abs(g["vini"][t]).max(axis=1, keepdims=True))
        assert (np.abs(x - g["vini"][t]) / sc).max() < RTOL, t
        assert np.abs(o14 - g["out14"][t]).max() < RTOL * max(1.0, np.abs(g["out14"][t]).max()), t
        assert np.abs(theta - g["theta"][t]).max() < RTOL, t
