"""40 Hz -> 100 Hz reference interpolation (PRMPCClass::XGetSolution_position_mod3,
RT/src/FastMPC/PRMPCClass.cpp:1170-1261).  The host part of the library (the 4x4 inverse, _t_end_footstep) is pinned
on the CPU against the reference's golden vectors, bit for bit.  The device kernel is held to the north star's
contract, 1e-9 relative: the reference evaluates the monomials with glibc pow (< 1 ulp, not always correctly rounded),
the kernel with a correctly rounded integer power, so single values may differ in the last bit.  The reference only
ever passes walktime = count_inteplotation in 1..n_t_int = 2 (gait_fast.cpp:113-131,487); the wide 0..2000 range of
the golden set extrapolates the cubic to |values| ~ 1e5 and is compared relative to each item's magnitude."""
import ctypes

import numpy as np
import pytest

import quadrupedal_loco_b200 as q
from tests.test_oracle_vs_ref import _oracle_interp, load


def test_ref_interp_model_matches_reference_golden():
    lib = q.load_library()
    g = load("interp_ref.npz")
    inv = np.zeros(16); te = ctypes.c_int(-1)
    assert lib.go1mpc_ref_interp_model(None, inv.ctypes.data, ctypes.byref(te)) == 0
    np.testing.assert_array_equal(inv, g["inv"])          # _AAA_inv_mod of the unmodified class, bit for bit
    assert te.value == int(g["t_end"][0]) == 1610
    assert lib.go1mpc_ref_interp_model(None, None, None) != 0


@pytest.mark.gpu
def test_gpu_ref_interp_vs_golden_and_oracle(mpc, oracle):
    import torch
    from tests.golden.make_golden import interp_inputs
    dev = torch.device("cuda", 0)
    g = load("interp_ref.npz")
    t_end = int(g["t_end"][0])
    for src, nh in (("golden", 4), ("synth", 10), ("synth", 1)):
        d = {k: g[k] for k in ("samples", "walktime")} if src == "golden" else interp_inputs(5000, seed=30 + nh)
        N = len(d["walktime"])
        want = g["out"][:, :9 + 3 * (nh - 1)] if src == "golden" else _oracle_interp(oracle, d, nh, t_end)[1]
        samples = torch.from_numpy(np.array(d["samples"].reshape(N, 12).T, order="C", copy=True)).to(dev)
        wt = torch.from_numpy(np.ascontiguousarray(d["walktime"], np.int32)).to(dev)
        out = torch.full((9 + 3 * (nh - 1), N), np.nan, dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        mpc.ref_interp(N, nh, wt, 0.01, samples, out)
        mpc.synchronize()
        got = out.cpu().numpy().T
        scale = np.maximum(1.0, np.abs(want).max(axis=1, keepdims=True))
        err = np.abs(got - want) / scale
        assert np.isfinite(got).all() and err.max() < 1e-9, f"{src} nh={nh}: max rel err {err.max():.3e}"
        late = d["walktime"] > t_end
        assert (got[late] == 0).all()                       # beyond _t_end_footstep: exact zeros
        real = d["walktime"] <= 3                           # the reference's own range of count_inteplotation
        assert real.sum() > 20 and np.abs(got[real] - want[real]).max() < 1e-12, f"{src} nh={nh} (operating range)"
