"""Synthetic, seeded workloads for tests and bench.py.

All generators use numpy's counter-based Philox with the seed of the config they
serve, so the CPU oracle, the GPU path and the reference arm see identical inputs.
"""
import numpy as np

SEED_CFG2 = 0xB2000002
SEED_CFG3 = 0xB2000003
SEED_CFG4 = 0xB2000004
SEED_CFG5 = 0xB2000005

HALF_HIP = 0.12675


def default_tx(tstep=0.7, dt_slow=0.025, n=27):
    """_tx of PRMPCClass::Initialize (RT/src/FastMPC/PRMPCClass.cpp:174-178)."""
    tx = np.zeros(n)
    for i in range(1, n):
        tx[i] = tx[i - 1] + tstep
        tx[i] = np.round(tx[i] / dt_slow) * dt_slow - 0.00001
    return tx


def body_mpc_inputs(B, nh, seed=SEED_CFG2, scale=1.0, tick_lo=100, tick_hi=1700, ref_amp=0.25, theta_clip=None):
    """cfg2-style body-MPC batch: randomised body-angle state, tick and reference windows.

    Returns dict(tick [B] i32, tx [B,27], theta [B,4], bstate [B,4], x_warm [B,2nh], refs [B,9,nh]).
    theta ~ U(+-0.12 rad), theta_dot ~ U(+-1.0 rad/s) (times `scale`), zmp_ref = support-foot
    centre + U(+-0.01), comacc_z ~ U(+-2), and the commanded body inclination
    bodyangle_ref ~ U(+-ref_amp) (offset + slope over the window).  With the reference's weights
    (beta = 5e9 on angle tracking) the controller is near dead-beat, so a command inside the
    +-10 deg limit never activates a constraint (every solve ends after the unconstrained step);
    ref_amp = 0.25 rad lets about 60 % of the instances command an inclination beyond the limit
    or a move that saturates the torque bound, which is what exercises the active-set
    iteration (mean ~4 constraints added, up to 16 at nh = 10).  scale > 1.5 additionally
    starts some instances outside the angle limit (infeasible solves, status 2); theta_clip keeps the
    initial angles inside +-theta_clip (bench.py's cfg3: 2x perturbation with the state inside the
    reference's own +-10 deg = 0.1745 rad envelope, so every QP has a solution).
    """
    rng = np.random.Generator(np.random.Philox(seed))
    tick = rng.integers(tick_lo, tick_hi + 1, size=B).astype(np.int32)
    tx = np.tile(default_tx(), (B, 1))
    theta = np.empty((B, 4))
    theta[:, 0] = rng.uniform(-0.12, 0.12, B) * scale
    theta[:, 1] = rng.uniform(-1.0, 1.0, B) * scale
    theta[:, 2] = rng.uniform(-0.12, 0.12, B) * scale
    theta[:, 3] = rng.uniform(-1.0, 1.0, B) * scale
    if theta_clip is not None:
        theta[:, 0] = np.clip(theta[:, 0], -theta_clip, theta_clip)
        theta[:, 2] = np.clip(theta[:, 2], -theta_clip, theta_clip)
    bstate = theta + rng.uniform(-0.01, 0.01, (B, 4))
    x_warm = np.zeros((B, 2 * nh))
    refs = np.zeros((B, 9, nh))
    # feet: a walking stance around a forward-moving body; constant over the short window
    xc = rng.uniform(0.0, 1.0, B)
    step = rng.uniform(-0.05, 0.15, B)
    lf = np.stack([xc + 0.5 * step, np.full(B, HALF_HIP)], 1) + rng.uniform(-0.01, 0.01, (B, 2))
    rf = np.stack([xc - 0.5 * step, np.full(B, -HALF_HIP)], 1) + rng.uniform(-0.01, 0.01, (B, 2))
    refs[:, 4, :] = rf[:, 0:1]; refs[:, 5, :] = rf[:, 1:2]
    refs[:, 6, :] = lf[:, 0:1]; refs[:, 7, :] = lf[:, 1:2]
    # zmp reference: under a foot, with the velocity-command perturbation
    use_l = rng.integers(0, 2, B).astype(bool)
    zc = np.where(use_l[:, None], lf, rf)
    refs[:, 0, :] = zc[:, 0:1] + rng.uniform(-0.01, 0.01, (B, nh)) * scale
    refs[:, 1, :] = zc[:, 1:2] + rng.uniform(-0.01, 0.01, (B, nh)) * scale
    # body-angle reference: offset + slope over the window
    a0 = rng.uniform(-ref_amp, ref_amp, (B, 2)) * scale
    a1 = rng.uniform(-0.02, 0.02, (B, 2)) * scale
    ramp = np.linspace(0.0, 1.0, nh)[None, :]
    refs[:, 2, :] = a0[:, 0:1] + a1[:, 0:1] * ramp
    refs[:, 3, :] = a0[:, 1:2] + a1[:, 1:2] * ramp
    refs[:, 8, :] = rng.uniform(-2.0, 2.0, (B, nh)) * scale
    return dict(tick=tick, tx=tx, theta=theta, bstate=bstate, x_warm=x_warm, refs=refs)


def random_qp(B, n, p, m, seed=1, paired=False, dup=False, infeasible_frac=0.0):
    """Random strictly convex QPs, feasible around a random point (column-major per problem)."""
    rng = np.random.Generator(np.random.Philox(seed))
    G = np.empty((B, n * n)); g0 = rng.standard_normal((B, n)) * 3
    CE = np.zeros((B, max(1, n * p))); ce0 = np.zeros((B, max(1, p)))
    CI = np.empty((B, n * m)); ci0 = np.empty((B, m))
    for b in range(B):
        A = rng.standard_normal((n, n))
        Gm = A @ A.T + 0.1 * np.eye(n)
        G[b] = Gm.ravel(order="F")
        x0 = rng.standard_normal(n) * 0.1
        if p:
            CEm = rng.standard_normal((n, p))
            CE[b, :n * p] = CEm.ravel(order="F"); ce0[b, :p] = -(CEm.T @ x0)
        CIm = rng.standard_normal((n, m))
        c0 = -(CIm.T @ x0) + np.abs(rng.standard_normal(m)) * 0.3
        if paired:
            h = m // 2
            CIm[:, h:2 * h] = -CIm[:, :h]
            c0[h:2 * h] = (CIm[:, :h].T @ x0) + np.abs(rng.standard_normal(h)) * 0.3
        if dup and m >= 2:
            CIm[:, 1] = CIm[:, 0]; c0[1] = c0[0]
        if rng.uniform() < infeasible_frac and m >= 2:
            CIm[:, 1] = -CIm[:, 0]; c0[1] = -c0[0] - 1.0
        CI[b] = CIm.ravel(order="F"); ci0[b] = c0
    return dict(G=G, g0=g0, CE=CE, ce0=ce0, CI=CI, ci0=ci0)


def step_timing_inputs(B, base_state, seed=SEED_CFG2, dt=0.025, Wn=None, amp=1.0, push_x=1.0, push_y=1.0, p_hi=23):
    """cfg2-style step-timing batch: every instance is a planner somewhere
    in its walk, its CoM on the nominal LIPM orbit of that step plus a random push.

    base_state: the 202-double default planner state (step tables of FootStepInputs/Initialize).
    Per instance: support period p ~ U{3..22}, elapsed samples k_yu ~ U{0..24}, tick
    i = round(tx[p-1]/dt) + k_yu.  Nominal orbit relative to the support foot: start
    -L[p-2]/2, end +L[p-1]/2 after ts (the boundary conditions step_timing_opti_loop itself
    uses, NLPClass_sqp.cpp:896-901).  Push: position U(+-0.01) m, velocity U(+-0.08) m/s in x and
    U(+-0.006) m, U(+-0.05) m/s in y; velocity command: Lxx_ref[p-1] += U(+-0.02); all times
    `amp`.  amp = 1 leaves about 3/4 of the QPs feasible with 1..4 constraints active; larger
    pushes make the reference's own formulation infeasible (its CoM-acceleration rows conflict),
    which the status codes report.  The warm start is the reference point (Lxx, Lyy, cosh(wT),
    sinh(wT)) of the previous tick; the end-of-step velocity reference is re-derived from it.
    push_x / push_y scale the x / y pushes on top of `amp`; p_hi bounds the support period (exclusive).
    bench.py uses p_hi = 16 (the forward-walking steps: on the backward steps 16-22 of the reference's
    own footstep plan about 15 % of the nominal QPs are already infeasible without any push) with
    push_x = 0.4, push_y = 0.75 -- x pushes of +-0.004 m / +-0.032 m/s at amp = 1 -- which leaves
    > 99 % of the QPs feasible at amp = 1 (cfg2) and about 93 % at amp = 2 (cfg3).
    Returns tick [B] i32, state [B,202], inp [B,20] (instance-major; transpose for the SoA ABI).
    """
    import math
    if Wn is None:
        Wn = math.sqrt(9.8 / 0.309458)
    rng = np.random.Generator(np.random.Philox(seed))
    st = np.tile(np.asarray(base_state, dtype=np.float64), (B, 1))
    ar = np.arange(B)
    p = rng.integers(3, p_hi, B)
    k_yu = rng.integers(0, 25, B)
    ki = np.round(st[ar, 27 + p - 1] / dt).astype(np.int64)
    tick = (ki + k_yu).astype(np.int32)
    fx = st[ar, 54 + p - 1]; fy = st[ar, 81 + p - 1]
    T = st[ar, p - 1]
    t = k_yu * dt
    st[ar, 135 + p - 1] += amp * rng.uniform(-0.02, 0.02, B)
    Lx = st[ar, 135 + p - 1]; Ly = st[ar, 162 + p - 1]
    isx = -0.5 * st[ar, 135 + p - 2]; isy = -0.5 * st[ar, 162 + p - 2]
    visx = (0.5 * Lx - isx * np.cosh(Wn * T)) / (np.sinh(Wn * T) / Wn)
    visy = (0.5 * Ly - isy * np.cosh(Wn * T)) / (np.sinh(Wn * T) / Wn)
    cx = fx + isx * np.cosh(Wn * t) + visx / Wn * np.sinh(Wn * t) + amp * push_x * rng.uniform(-0.01, 0.01, B)
    cvx = Wn * isx * np.sinh(Wn * t) + visx * np.cosh(Wn * t) + amp * push_x * rng.uniform(-0.08, 0.08, B)
    cy = fy + isy * np.cosh(Wn * t) + visy / Wn * np.sinh(Wn * t) + amp * push_y * rng.uniform(-0.006, 0.006, B)
    cvy = Wn * isy * np.sinh(Wn * t) + visy * np.cosh(Wn * t) + amp * push_y * rng.uniform(-0.05, 0.05, B)
    st[:, 189] = cx; st[:, 190] = cvx; st[:, 191] = 0.0
    st[:, 192] = cy; st[:, 193] = cvy; st[:, 194] = 0.0
    Tk_prev = T - (k_yu - 1) * dt
    st[:, 195] = Lx; st[:, 196] = Ly; st[:, 197] = np.cosh(Wn * Tk_prev); st[:, 198] = np.sinh(Wn * Tk_prev)
    st[:, 199] = Wn * isx * np.sinh(Wn * T) + visx * np.cosh(Wn * T)
    st[:, 200] = Wn * isy * np.sinh(Wn * T) + visy * np.cosh(Wn * T)
    inp = np.zeros((B, 20))
    inp[:, 7] = -0.12675; inp[:, 9] = 0.12675
    inp[:, 10:13] = 0.309458
    # _bjx1 left by the previous tick: the period of time i*dt... the planner's own value is the period of (i)*dt + dt
    st[:, 201] = p
    return tick, st, inp
