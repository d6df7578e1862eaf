"""GPU parity of the generic dense batched QP (go1mpc_qp_solve_batch, through the C ABI)
against the CPU oracle: primal to 1e-9 relative, identical final active set (same order),
identical iteration counters and status.  Shapes are the ones the reference solves:
step-timing (4,1,24), body MPC (8,0,48)/(20,0,120)/(40,0,240), GRF (12,12,24)-like."""
import numpy as np
import pytest

from quadrupedal_loco_b200 import synth

pytestmark = pytest.mark.gpu

RTOL = 1e-9   # north_star: 1e-9 relative on primal variables


def gpu_solve(mpc, n, p, m, d, x0=None):
    B = d["G"].shape[0]
    x = np.zeros((B, n)) if x0 is None else np.array(x0, dtype=float)
    cost = np.zeros(B); active = np.full((B, m + p), -99, np.int32); nact = np.zeros(B, np.int32)
    iters = np.zeros((B, 6), np.int32); status = np.full(B, -7, np.int32)
    G = np.ascontiguousarray(d["G"]); g0 = np.ascontiguousarray(d["g0"])
    CE = np.ascontiguousarray(d["CE"][:, :n * p]) if p else None
    ce0 = np.ascontiguousarray(d["ce0"][:, :p]) if p else None
    CI = np.ascontiguousarray(d["CI"]); ci0 = np.ascontiguousarray(d["ci0"])
    mpc.qp_solve_host(n, p, m, B, G, g0, CE, ce0, CI, ci0, x, cost, active, nact, iters, status)
    return dict(x=x, cost=cost, active=active, nactive=nact, iters=iters, status=status)


def assert_parity(g, o, n, label=""):
    assert np.array_equal(g["status"], o["status"]), f"{label}: status differs at {np.nonzero(g['status'] != o['status'])[0][:10]}"
    ok = o["status"] == 0
    assert np.array_equal(g["nactive"][ok], o["nactive"][ok]), f"{label}: active-set size differs"
    for b in np.nonzero(ok)[0]:
        k = o["nactive"][b]
        assert np.array_equal(g["active"][b, :k], o["active"][b, :k]), f"{label}: active set differs at problem {b}"
    assert np.array_equal(g["iters"][ok][:, :4], o["iters"][ok]), f"{label}: iteration counters differ"
    scale = np.maximum(1.0, np.abs(o["x"]).max(axis=1, keepdims=True))
    err = (np.abs(g["x"] - o["x"]) / scale)[ok]
    assert err.max(initial=0.0) < RTOL, f"{label}: primal rel err {err.max():.3e}"
    cs = np.maximum(1.0, np.abs(o["cost"][ok]))
    assert (np.abs(g["cost"][ok] - o["cost"][ok]) / cs).max(initial=0.0) < 1e-8, f"{label}: cost"
    bad = ~ok
    assert np.array_equal(np.isinf(g["cost"][bad]), np.isinf(o["cost"][bad]))


@pytest.mark.parametrize("shape", [(4, 1, 24), (8, 0, 48), (20, 0, 120), (12, 2, 24), (40, 0, 240), (2, 0, 3), (33, 3, 70)])
def test_random_paired(mpc, oracle, shape):
    n, p, m = shape
    B = 192 if n <= 20 else 48
    d = synth.random_qp(B, n, p, m, seed=100 + n, paired=True)
    assert_parity(gpu_solve(mpc, n, p, m, d), oracle.qp_solve_batch(n, p, m, d), n, str(shape))


def test_duplicate_and_infeasible(mpc, oracle):
    n, p, m = 6, 0, 14
    d = synth.random_qp(128, n, p, m, seed=77, dup=True, infeasible_frac=0.3)
    g = gpu_solve(mpc, n, p, m, d); o = oracle.qp_solve_batch(n, p, m, d)
    assert (o["status"] == 2).any() and (o["status"] == 0).any()
    assert_parity(g, o, n, "dup/infeasible")


def test_not_pd_leaves_x(mpc, oracle):
    n, m = 4, 6
    d = synth.random_qp(16, n, 0, m, seed=5)
    Gm = np.eye(n); Gm[0, 0] = -1.0
    d["G"][::2] = Gm.ravel(order="F")
    x0 = np.arange(16 * n, dtype=float).reshape(16, n)
    g = gpu_solve(mpc, n, 0, m, d, x0=x0); o = oracle.qp_solve_batch(n, 0, m, d, x0=x0)
    assert (g["status"][::2] == 1).all() and np.isinf(g["cost"][::2]).all()
    np.testing.assert_array_equal(g["x"][::2], x0[::2])
    assert_parity(g, o, n, "not-pd")


def test_quadprogpp_demo(mpc):
    d = dict(G=np.array([[4., -2, -2, 4]]), g0=np.array([[6., 0]]), CE=np.array([[1., 1]]), ce0=np.array([[-3.]]),
             CI=np.array([[1., 0, 0, 1, 1, 1]]), ci0=np.array([[0., 0, -2]]))
    g = gpu_solve(mpc, 2, 1, 3, d)
    np.testing.assert_allclose(g["x"][0], [1, 2], atol=1e-12)
    assert abs(g["cost"][0] - 12) < 1e-12 and g["status"][0] == 0 and g["active"][0, 0] == -1 and g["nactive"][0] == 1


def test_zero_equality_columns_are_skipped(mpc, oracle):
    """GRF-style: all-zero CE columns are skipped by the reference (EiQuadProg.cpp:240) but me = p."""
    n, p, m = 6, 3, 8
    d = synth.random_qp(64, n, p, m, seed=9)
    CE = d["CE"].reshape(64, p, n); CE[:, 1, :] = 0.0; d["ce0"][:, 1] = 0.0
    d["CE"] = CE.reshape(64, n * p)
    assert_parity(gpu_solve(mpc, n, p, m, d), oracle.qp_solve_batch(n, p, m, d), n, "zero-CE")


def test_ragged_batch_sizes(mpc, oracle):
    n, p, m = 8, 0, 48
    for B in (1, 3, 5, 1000):
        d = synth.random_qp(B, n, p, m, seed=B, paired=True)
        assert_parity(gpu_solve(mpc, n, p, m, d), oracle.qp_solve_batch(n, p, m, d), n, f"B={B}")
    # empty batch is a no-op
    mpc.qp_solve_host(n, p, m, 0, np.zeros(1), np.zeros(1), None, None, np.zeros(1), np.zeros(1), np.zeros(1))


def test_device_pointer_entry(mpc, oracle):
    import torch
    n, p, m, B = 8, 0, 48, 256
    d = synth.random_qp(B, n, p, m, seed=21, paired=True)
    dev = torch.device("cuda", 0)
    t = {k: torch.from_numpy(np.ascontiguousarray(d[k])).to(dev) for k in ("G", "g0", "CI", "ci0")}
    x = torch.zeros(B, n, dtype=torch.float64, device=dev); cost = torch.zeros(B, dtype=torch.float64, device=dev)
    act = torch.zeros(B, m, dtype=torch.int32, device=dev); na = torch.zeros(B, dtype=torch.int32, device=dev)
    it = torch.zeros(B, 6, dtype=torch.int32, device=dev); st = torch.zeros(B, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    before = mpc.launch_count
    mpc.qp_solve(n, p, m, B, t["G"], t["g0"], None, None, t["CI"], t["ci0"], x, cost, act, na, it, st)
    mpc.synchronize()
    assert mpc.launch_count == before + 1
    g = dict(x=x.cpu().numpy(), cost=cost.cpu().numpy(), active=act.cpu().numpy(), nactive=na.cpu().numpy(),
             iters=it.cpu().numpy(), status=st.cpu().numpy())
    assert_parity(g, oracle.qp_solve_batch(n, p, m, d), n, "device")


def test_bad_arguments(mpc):
    import quadrupedal_loco_b200 as q
    with pytest.raises(q.Go1MpcError):
        mpc.qp_solve_host(200, 0, 4, 1, np.zeros(40000), np.zeros(200), None, None, np.zeros(800), np.zeros(4), np.zeros(200))
