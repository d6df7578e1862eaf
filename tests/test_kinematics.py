"""Leg kinematics: the oracle against the reference's golden outputs (CPU) and the CUDA kernels
against the oracle and the golden outputs (GPU).  Reference: Kinematicclass,
GO1/src/kinematics/Kinematics.cpp:63-304; MATLAB demo inputs kinematics_matlab/*.m."""
import numpy as np
import pytest

from tests.test_oracle_vs_ref import load


def kin_inputs(N, seed=9):
    rng = np.random.Generator(np.random.Philox(seed))
    q = np.stack([rng.uniform(-0.6, 0.6, N), rng.uniform(0.2, 1.4, N), rng.uniform(-2.2, -0.9, N)], 1)
    bp = rng.uniform(-0.05, 0.05, (N, 3)) + [0, 0, 0.31]; br = rng.uniform(-0.25, 0.25, (N, 3))
    leg = rng.integers(0, 4, N).astype(np.int32)
    qini = q + rng.uniform(-0.15, 0.15, (N, 3))
    return q, bp, br, leg, qini


# ----------------------------------------------------------------------------- CPU
def test_oracle_kinematics_vs_reference_golden(oracle):
    g = load("kin_ref.npz")
    pos, J = oracle.leg_fk(g["q"], g["leg"])
    np.testing.assert_allclose(pos, g["fk"], rtol=0, atol=1e-12); np.testing.assert_allclose(J, g["fkJ"], rtol=0, atol=1e-12)
    pos, J = oracle.leg_fk(g["q"], g["leg"], g["bp"], g["br"])
    np.testing.assert_allclose(pos, g["fkg"], rtol=0, atol=1e-12); np.testing.assert_allclose(J, g["fkgJ"], rtol=0, atol=1e-12)
    q, J, it = oracle.leg_ik(g["fk"], g["qini"], g["leg"])
    np.testing.assert_allclose(q, g["ik"], rtol=0, atol=1e-12); np.testing.assert_allclose(J, g["ikJ"], rtol=0, atol=1e-12)
    q, J, it = oracle.leg_ik(g["fkg"], g["qini"], g["leg"], g["bp"], g["br"])
    np.testing.assert_allclose(q, g["ikg"], rtol=0, atol=1e-12); np.testing.assert_allclose(J, g["ikgJ"], rtol=0, atol=1e-12)


def test_oracle_jacobian_is_the_derivative(oracle):
    q, bp, br, leg, _ = kin_inputs(40)
    for glob in (False, True):
        a = (bp, br) if glob else (None, None)
        p0, J = oracle.leg_fk(q, leg, *a)
        for k in range(3):
            h = 1e-6
            dq = np.zeros(3); dq[k] = h
            pp, _ = oracle.leg_fk(q + dq, leg, *a); pm, _ = oracle.leg_fk(q - dq, leg, *a)
            np.testing.assert_allclose((pp - pm) / (2 * h), J.reshape(-1, 3, 3)[:, :, k], atol=2e-9)


def test_oracle_ik_round_trip_and_matlab_demo(oracle):
    q, bp, br, leg, qini = kin_inputs(64)
    pg, _ = oracle.leg_fk(q, leg, bp, br)
    qs, _, it = oracle.leg_ik(pg, qini, leg, bp, br)
    back, _ = oracle.leg_fk(qs, leg, bp, br)
    assert np.abs(back - pg).max() < 1.1e-3 and (it <= 15).all() and (it > 0).any()    # stops at |dp|^2 <= 1e-6
    # kinematics_matlab/forward_kin_go1.m, inverse_kin_go1.m: q = (0, 0.6, -1) -> foot, target [0.1881, -0.12765, -0.32]
    p, _ = oracle.leg_fk(np.array([[0, 0.6, -1.0]]), [0])
    assert abs(p[0, 1] - (-0.04675 - 0.08)) < 1e-15 and p[0, 2] < -0.3
    qd, _, it = oracle.leg_ik(np.array([[0.1881, -0.12765, -0.32]]), np.array([[0, 0.6, -1.0]]), [0])
    assert np.isfinite(qd).all()


# ----------------------------------------------------------------------------- GPU
def _soa(a):
    return np.array(np.asarray(a, dtype=np.float64).T, order="C", copy=True)


def gpu_fk(mpc, q, leg, bp=None, br=None, device=False):
    B = len(leg)
    pos = np.zeros((3, B)); J = np.zeros((9, B))
    args = (_soa(q), np.ascontiguousarray(leg, np.int32), None if bp is None else _soa(bp), None if br is None else _soa(br))
    if device:
        import torch
        dev = torch.device("cuda", 0)
        t = [None if a is None else torch.from_numpy(a).to(dev) for a in args]
        tp = torch.zeros(3, B, dtype=torch.float64, device=dev); tj = torch.zeros(9, B, dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        mpc.leg_fk(B, t[0], t[1], t[2], t[3], tp, tj); mpc.synchronize()
        return tp.cpu().numpy().T.copy(), tj.cpu().numpy().T.copy()
    mpc.leg_fk_host(B, args[0], args[1], args[2], args[3], pos, J)
    return pos.T.copy(), J.T.copy()


def gpu_ik(mpc, pdes, qini, leg, bp=None, br=None):
    B = len(leg)
    q = np.zeros((3, B)); J = np.zeros((9, B)); it = np.zeros(B, np.int32)
    mpc.leg_ik_host(B, _soa(pdes), _soa(qini), np.ascontiguousarray(leg, np.int32), None if bp is None else _soa(bp),
                    None if br is None else _soa(br), q, J, it)
    return q.T.copy(), J.T.copy(), it


@pytest.mark.gpu
def test_gpu_kinematics_vs_reference_golden(mpc):
    g = load("kin_ref.npz")
    pos, J = gpu_fk(mpc, g["q"], g["leg"])
    np.testing.assert_allclose(pos, g["fk"], rtol=0, atol=1e-12); np.testing.assert_allclose(J, g["fkJ"], rtol=0, atol=1e-12)
    pos, J = gpu_fk(mpc, g["q"], g["leg"], g["bp"], g["br"], device=True)
    np.testing.assert_allclose(pos, g["fkg"], rtol=0, atol=1e-12); np.testing.assert_allclose(J, g["fkgJ"], rtol=0, atol=1e-12)
    q, J, it = gpu_ik(mpc, g["fk"], g["qini"], g["leg"])
    np.testing.assert_allclose(q, g["ik"], rtol=0, atol=1e-9); np.testing.assert_allclose(J, g["ikJ"], rtol=0, atol=1e-9)
    q, J, it = gpu_ik(mpc, g["fkg"], g["qini"], g["leg"], g["bp"], g["br"])
    np.testing.assert_allclose(q, g["ikg"], rtol=0, atol=1e-9); np.testing.assert_allclose(J, g["ikgJ"], rtol=0, atol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("B", [1, 255, 4096, 100000])
def test_gpu_kinematics_vs_oracle(mpc, oracle, B):
    q, bp, br, leg, qini = kin_inputs(B, seed=B)
    n = min(B, 3000)          # the python-loop oracle is the slow side
    pos, J = gpu_fk(mpc, q, leg)
    po, Jo = oracle.leg_fk(q[:n], leg[:n])
    np.testing.assert_allclose(pos[:n], po, rtol=0, atol=1e-13); np.testing.assert_allclose(J[:n], Jo, rtol=0, atol=1e-13)
    posg, Jg = gpu_fk(mpc, q, leg, bp, br)
    po, Jo = oracle.leg_fk(q[:n], leg[:n], bp[:n], br[:n])
    np.testing.assert_allclose(posg[:n], po, rtol=0, atol=1e-13); np.testing.assert_allclose(Jg[:n], Jo, rtol=0, atol=1e-13)
    qs, Js, it = gpu_ik(mpc, posg, qini, leg, bp, br)
    qo, Jo, ito = oracle.leg_ik(posg[:n], qini[:n], leg[:n], bp[:n], br[:n])
    assert np.array_equal(it[:n], ito), "IK update counts differ"
    np.testing.assert_allclose(qs[:n], qo, rtol=0, atol=1e-9)
    ql, Jl, itl = gpu_ik(mpc, pos, qini, leg)
    qo, Jo, ito = oracle.leg_ik(pos[:n], qini[:n], leg[:n])
    assert np.array_equal(itl[:n], ito)
    np.testing.assert_allclose(ql[:n], qo, rtol=0, atol=1e-9)
    # size-independent property on the whole batch: FK(IK_g(p)) lands within the stop radius of p
    back, _ = gpu_fk(mpc, qs, leg, bp, br)
    conv = it < 15
    assert conv.mean() > 0.95 and np.abs(back - posg)[conv].max() < 1.1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("gait_mode", [101, 102, 103])
def test_gpu_servo_kin_tick_vs_oracle(mpc, oracle, gait_mode):
    """servo.cpp:935-1051: gait_mode leg mapping + 4 x Inverse_kinematics_g from the previous joint angles,
    against the same mapping done in numpy + the oracle IK."""
    import torch
    B = 3000
    rng = np.random.Generator(np.random.Philox(gait_mode))
    HW = 0.12675
    com = rng.uniform(-0.03, 0.03, (B, 3)) + [0, 0, 0.3]; theta = rng.uniform(-0.15, 0.15, (B, 3))
    rfoot = rng.uniform(-0.04, 0.04, (B, 3)); rfoot[:, 1] -= HW; rfoot[:, 2] = np.abs(rfoot[:, 2]) * 0.5
    lfoot = rng.uniform(-0.04, 0.04, (B, 3)); lfoot[:, 1] += HW; lfoot[:, 2] = np.abs(lfoot[:, 2]) * 0.5
    qh = np.array([0, 0.87, -1.5])                       # homing pose, servo.cpp:767
    legs = np.arange(4)
    hom, _ = oracle.leg_fk(np.tile(qh, (4, 1)), legs, np.tile([0, 0, 0.3], (4, 1)), np.zeros((4, 3)))
    homing = np.tile(hom.reshape(1, 12), (B, 1))
    q0 = np.tile(qh, (B, 4)) + rng.uniform(-0.05, 0.05, (B, 12))
    y_off = 0.85
    dev = torch.device("cuda", 0)
    soa = lambda a: torch.from_numpy(np.array(a.T, order="C", copy=True)).to(dev)
    tq = soa(q0); tj = torch.zeros(36, B, dtype=torch.float64, device=dev); tf = torch.zeros(12, B, dtype=torch.float64, device=dev)
    ti = torch.zeros(4, B, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    mpc.servo_kin_tick(B, gait_mode, y_off, soa(com), soa(theta), soa(rfoot), soa(lfoot), soa(homing), tq, tj, tf, ti)
    mpc.synchronize()
    gq, gj, gf, gi = tq.cpu().numpy().T, tj.cpu().numpy().T, tf.cpu().numpy().T, ti.cpu().numpy().T
    right = {101: [1, 0, 1, 0], 102: [1, 0, 0, 1], 103: [1, 1, 0, 0]}[gait_mode]      # FR, FL, RR, RL follow the right foot?
    bp = com.copy(); bp[:, 1] *= y_off
    n = 600
    for leg in range(4):
        vf = rfoot if right[leg] else lfoot
        pdes = homing[:, 3 * leg:3 * leg + 3] + vf
        pdes[:, 1] += HW if right[leg] else -HW
        np.testing.assert_allclose(gf[:, 3 * leg:3 * leg + 3], pdes, rtol=0, atol=1e-15)
        qo, Jo, ito = oracle.leg_ik(pdes[:n], q0[:n, 3 * leg:3 * leg + 3], np.full(n, leg, np.int32), bp[:n], theta[:n])
        assert np.array_equal(gi[:n, leg], ito)
        np.testing.assert_allclose(gq[:n, 3 * leg:3 * leg + 3], qo, rtol=0, atol=1e-9)
        np.testing.assert_allclose(gj[:n, 9 * leg:9 * leg + 9], Jo, rtol=0, atol=1e-9)
