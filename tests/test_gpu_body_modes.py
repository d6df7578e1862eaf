"""The device paths of the body-inclination MPC tick at horizon 10 -- the combined warp-per-instance
kernel ("fast"), roll / pitch halves side by side in one warp ("split"), the three-launch path with
register-resident solver state ("tri") and the any-horizon interleaved-halves kernel ("duo", body_duo.cu: what every
horizon other than 4 / 10 runs) -- each forced through GO1MPC_BODY_MODE and checked against the CPU
oracle of PRMPCClass::body_theta_mpc (RT/src/FastMPC/PRMPCClass.cpp:379-714) exactly like test_gpu_body.py:
primal 1e-9 relative, identical ORDERED final active set and iteration counters (the split paths rebuild the
reference's interleaving of the two halves from their logs), bit-exact phase indices.  Instances the split
paths cannot reproduce (infeasible / degenerate halves) must come back through the combined kernel."""
import os

import numpy as np
import pytest

import quadrupedal_loco_b200 as q
from quadrupedal_loco_b200 import synth
from tests.test_gpu_body import run_gpu, run_oracle, assert_body_parity

pytestmark = pytest.mark.gpu

# "split" (body_split.cu) is an A/B variant: compiled into the library only with GO1MPC_BUILD_AB=1
MODES = ["fast", "tri", "duo"] + (["split"] if os.environ.get("GO1MPC_BUILD_AB") == "1" else [])


@pytest.fixture(params=MODES)
def handle(request):
    old = os.environ.get("GO1MPC_BODY_MODE")
    os.environ["GO1MPC_BODY_MODE"] = request.param
    h = q.Go1Mpc(0)
    try:
        yield request.param, h
    finally:
        h.close()
        if old is None:
            os.environ.pop("GO1MPC_BODY_MODE", None)
        else:
            os.environ["GO1MPC_BODY_MODE"] = old


@pytest.mark.parametrize("B", [1, 3, 33, 777])
def test_modes_parity_ragged_batches(handle, oracle, B):
    mode, h = handle
    d = synth.body_mpc_inputs(B, 10, seed=synth.SEED_CFG2 + 17 * B)
    out, diag = run_gpu(h, 10, d, device=(B % 2 == 1))
    r = run_oracle(oracle, 10, d)
    assert_body_parity(out, diag, r, 10, f"{mode} B={B}")
    assert h.body_guard_trips() == 0


def test_modes_parity_large_perturbation_and_handover(handle, oracle):
    """cfg3-style 2x perturbation: drops, degenerate adds, infeasible instances.  The split paths hand those
    they cannot reproduce to the combined kernel; the results are the oracle's either way."""
    mode, h = handle
    nh, B = 10, 3000
    d = synth.body_mpc_inputs(B, nh, seed=synth.SEED_CFG3, scale=2.0)
    h0 = h.body_handover_total()
    out, diag = run_gpu(h, nh, d)
    r = run_oracle(oracle, nh, d)
    assert_body_parity(out, diag, r, nh, f"{mode} cfg3")
    handed = h.body_handover_total() - h0
    hard = int((r["status"] != 0).sum())
    assert hard > 0, "workload has no infeasible instance"
    if mode == "fast":
        assert handed == 0
    elif mode == "duo":
        assert handed <= B // 50           # body_duo.cu solves infeasible instances itself; it hands over one rare corner only
    else:
        assert handed >= hard, (handed, hard)          # every non-converged instance went through the combined kernel
        assert handed <= hard + B // 50, (handed, hard)  # ... and (almost) only those
    assert h.body_guard_trips() == 0


def test_modes_gated_ticks_and_stale_outputs(handle, oracle):
    mode, h = handle
    nh, B = 10, 200
    d = synth.body_mpc_inputs(B, nh, seed=7)
    d["tick"][::3] = np.arange(len(d["tick"][::3])) % 100      # i < 100: the reference returns its stale members
    d["x_warm"][:] = 0.125
    prev = np.random.default_rng(1).standard_normal((B, 14))
    out, diag = run_gpu(h, nh, d, out14_prev=prev)
    r = run_oracle(oracle, nh, d, out14_prev=prev)
    assert_body_parity(out, diag, r, nh, f"{mode} gated mix")
    g = np.zeros(B, bool); g[::3] = True
    np.testing.assert_array_equal(out[g, :14], prev[g])
    np.testing.assert_array_equal(out[g, 14:18], d["theta"][g])
    np.testing.assert_array_equal(out[g, 18:18 + 2 * nh], d["x_warm"][g])
    assert (diag[g, 0] == -1).all()


def test_modes_closed_loop(handle, oracle):
    """12 closed-loop ticks, state carried on both sides, alternating host / device entry."""
    mode, h = handle
    nh, B, T = 10, 96, 12
    d = synth.body_mpc_inputs(B, nh, seed=23)
    d["tick"][:] = 150 + (np.arange(B) % 60)
    go = dict(d); oo = dict(d)
    g14 = np.zeros((B, 14)); o14 = np.zeros((B, 14))
    for t in range(T):
        out, diag = run_gpu(h, nh, go, out14_prev=g14, device=(t % 2 == 0))
        r = run_oracle(oracle, nh, oo, out14_prev=o14)
        assert_body_parity(out, diag, r, nh, f"{mode} tick+{t}")
        g14 = out[:, :14].copy(); o14 = r["out14"]
        go = dict(go, theta=out[:, 14:18].copy(), x_warm=out[:, 18:18 + 2 * nh].copy(), tick=go["tick"] + 1)
        oo = dict(oo, theta=r["theta"], x_warm=r["x"], tick=oo["tick"] + 1)


def test_modes_agree_with_each_other_at_full_size(oracle):
    """cfg2 at 4096 and 65536 instances: the three paths give the same diagnostics (status, ordered active set,
    counters, algorithmic flop count) for every instance and the same primal to 1e-12."""
    nh = 10
    for B in (4096, 65536):
        d = synth.body_mpc_inputs(B, nh, seed=synth.SEED_CFG2)
        res = {}
        for mode in MODES:
            os.environ["GO1MPC_BODY_MODE"] = mode
            h = q.Go1Mpc(0)
            try:
                res[mode] = run_gpu(h, nh, d)
                assert h.body_guard_trips() == 0
            finally:
                h.close()
                os.environ.pop("GO1MPC_BODY_MODE", None)
        for mode in MODES[1:]:
            np.testing.assert_array_equal(res[mode][1], res["fast"][1], err_msg=f"diagnostics {mode} vs fast, B={B}")
            x0, x1 = res["fast"][0], res[mode][0]
            err = np.abs(x1 - x0) / np.maximum(1.0, np.abs(x0).max(axis=1, keepdims=True))
            assert err.max() < 1e-12, (mode, B, err.max())


def test_tri_handover_list_overflow():
    """More hand-overs than the list holds (8190): the combined kernel then redoes the whole batch.  3x perturbation,
    40000 instances, ~55 % non-converged: the three-launch path must still equal the combined-solve kernel."""
    nh, B = 10, 40000
    d = synth.body_mpc_inputs(B, nh, seed=108, scale=3.0)
    res = {}
    for mode in ("fast", "tri"):
        os.environ["GO1MPC_BODY_MODE"] = mode
        h = q.Go1Mpc(0)
        try:
            h0 = h.body_handover_total()
            res[mode] = run_gpu(h, nh, d, device=True)
            res[mode + "_handed"] = h.body_handover_total() - h0
            assert h.body_guard_trips() == 0
        finally:
            h.close()
            os.environ.pop("GO1MPC_BODY_MODE", None)
    assert res["tri_handed"] > 8190, res["tri_handed"]
    da, db = res["fast"][1], res["tri"][1]
    conv = da[:, 0] == 0
    assert conv.sum() > B // 4
    np.testing.assert_array_equal(da[:, 0], db[:, 0])
    np.testing.assert_array_equal(da[conv], db[conv])
    xa, xb = res["fast"][0], res["tri"][0]
    err = np.abs(xa - xb) / np.maximum(1.0, np.abs(xa).max(axis=1, keepdims=True))
    err = np.where(np.isfinite(err), err, 0.0)
    assert err.max() < 1e-9


@pytest.mark.parametrize("B,scale", [(1, 1.0), (257, 1.0), (3000, 1.0), (3000, 2.0)])
def test_tri_at_the_reference_horizon(oracle, B, scale):
    """nh = 4 (PRMPCClass.h:34, the horizon the reference ships) through the three-launch path."""
    nh = 4
    os.environ["GO1MPC_BODY_MODE"] = "tri"
    h = q.Go1Mpc(0)
    try:
        d = synth.body_mpc_inputs(B, nh, seed=synth.SEED_CFG2 + B, scale=scale)
        l0 = h.launch_count
        out, diag = run_gpu(h, nh, d, device=True)
        assert h.launch_count - l0 == 4          # setup, solve, merge, list-mode combined kernel
        r = run_oracle(oracle, nh, d)
        assert_body_parity(out, diag, r, nh, f"tri nh=4 B={B} scale={scale}")
        assert h.body_guard_trips() == 0
        if B >= 257:
            assert (r["nactive"] > 0).any()
    finally:
        h.close()
        os.environ.pop("GO1MPC_BODY_MODE", None)


def test_two_devices_in_one_process(oracle):
    """Handles on two GPUs of one process: the kernels' opt-in shared-memory attributes and occupancy caches are per
    device.  Skipped on a single-GPU box."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    nh, B = 10, 2500
    d = synth.body_mpc_inputs(B, nh, seed=77)
    r = run_oracle(oracle, nh, d)
    for dev in (0, 1):
        for mode in ("tri", "fast"):
            os.environ["GO1MPC_BODY_MODE"] = mode
            h = q.Go1Mpc(dev)
            try:
                out, diag = run_gpu(h, nh, d)      # host entry: the handle's own device
                assert_body_parity(out, diag, r, nh, f"device {dev} {mode}")
            finally:
                h.close()
                os.environ.pop("GO1MPC_BODY_MODE", None)


@pytest.mark.parametrize("nh,B", [(20, 1024), (40, 256), (33, 300), (7, 500)])
def test_duo_horizons_large_perturbation(mpc, oracle, nh, B):
    """Horizons other than 4 / 10 run body_duo.cu (cfg4: 20, 40): 2x perturbation -- drops, infeasible instances -- against
    the oracle, ordered active set and counters included; the hand-over list stays (almost) empty."""
    d = synth.body_mpc_inputs(B, nh, seed=synth.SEED_CFG3 + nh, scale=2.0)
    h0 = mpc.body_handover_total()
    out, diag = run_gpu(mpc, nh, d)
    r = run_oracle(oracle, nh, d)
    assert_body_parity(out, diag, r, nh, f"duo nh={nh}")
    assert (r["iters"][:, 2] > 0).any(), "workload never drops a constraint"
    assert mpc.body_handover_total() - h0 <= B // 50


def test_dense_kernel_alone_still_matches(oracle):
    """GO1MPC_FORCE_GENERIC=1: the dense warp-per-problem kernel (body_mpc.cu) on its own -- it is also the list-mode
    fallback behind body_duo.cu."""
    old = os.environ.get("GO1MPC_FORCE_GENERIC")
    os.environ["GO1MPC_FORCE_GENERIC"] = "1"
    h = q.Go1Mpc(0)
    try:
        for nh, B in ((16, 200), (10, 300)):
            d = synth.body_mpc_inputs(B, nh, seed=synth.SEED_CFG3 + 3 * nh, scale=2.0)
            out, diag = run_gpu(h, nh, d)
            assert_body_parity(out, diag, run_oracle(oracle, nh, d), nh, f"dense nh={nh}")
    finally:
        h.close()
        if old is None:
            os.environ.pop("GO1MPC_FORCE_GENERIC", None)
        else:
            os.environ["GO1MPC_FORCE_GENERIC"] = old
