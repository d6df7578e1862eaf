"""The rule the roll / pitch split kernels (body_split.cu, body_tri.cu) rest on, checked on the CPU against the
oracle restatement of Eigen::QP::solve_quadprog2 (RT/src/utils/EiQuadProg/EiQuadProg.cpp:172-491):

  for G = blockdiag(H, H) and inequality columns that each touch ONE half, the reference's n-variable solve is an
  interleaving of the two independent n/2-variable solves; which half moves next is decided at step 1 / step 2
  (cpp:282-342) by the most negative slack over both halves, lowest constraint index on ties.

Random block-structured QPs with the body MPC's column layout (blocks 0,1 -> half 0; 2,3 -> half 1; 4,5 -> half 0;
6,7 -> half 1).  Checked: primal, cost, the active SET and the counters always; the ORDER of the active set for
every problem whose halves dropped nothing (then a half's selection slacks can be recomputed from its own ordered
active set, which is what the kernels log on the fly)."""
import numpy as np


def make_problem(rng, k):
    """k variables per half; 8 blocks of k constraint columns, rows of a lower-triangular P (angle-like) and
    single-variable bounds (torque-like)."""
    n = 2 * k
    A = rng.standard_normal((k, k))
    H = A @ A.T + k * np.eye(k)
    G = np.zeros((n, n)); G[:k, :k] = H; G[k:, k:] = H
    g0 = rng.standard_normal(n) * 6.0
    Pm = np.tril(rng.uniform(0.2, 1.0, (k, k)))
    lim_a = rng.uniform(0.3, 0.8); lim_t = rng.uniform(0.4, 1.0)
    m = 8 * k
    CI = np.zeros((n, m)); ci0 = np.zeros(m)
    for blk in range(8):
        half = (blk >> 1) & 1
        sgn = 1.0 if (blk & 1) else -1.0
        for j in range(k):
            col = blk * k + j
            if blk < 4:
                CI[half * k:(half + 1) * k, col] = sgn * Pm[j]
                ci0[col] = lim_a + rng.uniform(-0.05, 0.05)
            else:
                CI[half * k + j, col] = sgn
                ci0[col] = lim_t
    return G, g0, CI, ci0


def half_of(col, k):
    return ((col // k) >> 1) & 1


def half_problem(G, g0, CI, ci0, k, h):
    cols = [c for c in range(CI.shape[1]) if half_of(c, k) == h]
    sl = slice(h * k, (h + 1) * k)
    return G[sl, sl], g0[sl], CI[sl][:, cols], ci0[cols], cols


def x_with_active(Hh, gh, Ch, ch, act):
    """minimiser of the half's objective with the listed columns active (as equalities)"""
    k = len(gh)
    if not act:
        return -np.linalg.solve(Hh, gh)
    N = Ch[:, act]
    K = np.block([[Hh, -N], [N.T, np.zeros((len(act), len(act)))]])
    rhs = np.concatenate([-gh, -ch[act]])
    return np.linalg.solve(K, rhs)[:k]


def test_combined_solve_is_an_interleaving_of_the_half_solves(oracle):
    rng = np.random.Generator(np.random.Philox(20261018))
    k = 5
    n, m = 2 * k, 8 * k
    checked_order = 0; with_both = 0
    for trial in range(150):
        G, g0, CI, ci0 = make_problem(rng, k)
        F = lambda a: np.asfortranarray(a).ravel(order="F")
        rc = oracle.qp_solve(n, 0, m, F(G), g0, None, None, F(CI), ci0)
        if rc["status"] != 0:
            continue
        halves = []
        for h in range(2):
            Hh, gh, Ch, ch, cols = half_problem(G, g0, CI, ci0, k, h)
            rh = oracle.qp_solve(k, 0, len(cols), F(Hh), gh, None, None, F(Ch), ch)
            assert rh["status"] == 0
            halves.append((rh, Hh, gh, Ch, ch, cols))
        xh = np.concatenate([halves[0][0]["x"], halves[1][0]["x"]])
        np.testing.assert_allclose(rc["x"], xh, rtol=0, atol=1e-10)
        assert abs(rc["cost"] - (halves[0][0]["cost"] + halves[1][0]["cost"])) < 1e-9 * max(1.0, abs(rc["cost"]))
        glob = [[halves[h][5][a] for a in halves[h][0]["active"]] for h in range(2)]
        assert sorted(rc["active"].tolist()) == sorted(glob[0] + glob[1])
        it_c = rc["iters"]; it0, it1 = halves[0][0]["iters"], halves[1][0]["iters"]
        assert it_c[0] == it0[0] + it1[0] - 1            # step-1 passes: the final one is shared
        assert it_c[1] == it0[1] + it1[1] and it_c[2] == it0[2] + it1[2] and it_c[3] == 0 == it0[3] + it1[3]
        if it0[2] or it1[2]:
            continue                                      # a drop: the order needs the per-pass log the kernels keep
        # selection slacks of each half from its own ordered active set, then the merge rule
        events = []
        for h in range(2):
            rh, Hh, gh, Ch, ch, cols = halves[h]
            act = rh["active"].tolist()
            for j, a in enumerate(act):
                x = x_with_active(Hh, gh, Ch, ch, act[:j])
                events.append((h, j, float(Ch[:, a] @ x + ch[a]), cols[a]))
        order = []
        ptr = [0, 0]
        per_half = [[e for e in events if e[0] == h] for h in range(2)]
        while ptr[0] < len(per_half[0]) or ptr[1] < len(per_half[1]):
            cand = [per_half[h][ptr[h]] for h in range(2) if ptr[h] < len(per_half[h])]
            pick = min(cand, key=lambda e: (e[2], e[3]))   # most negative slack, lowest index on ties
            order.append(pick[3]); ptr[pick[0]] += 1
        assert order == rc["active"].tolist(), (trial, order, rc["active"].tolist())
        checked_order += 1
        with_both += bool(glob[0] and glob[1])
    assert checked_order >= 60 and with_both >= 20, (checked_order, with_both)
