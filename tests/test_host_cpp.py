"""The C++ mirror of the reference classes (quadrupedal_loco_b200/host/go1mpc.hpp): it compiles
headless with g++ against the C ABI only (CPU check), and -- on a GPU -- the calls made the way
the reference's callers make them reproduce the CPU oracle."""
import json
import os
import subprocess

import numpy as np
import pytest

import quadrupedal_loco_b200 as q

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_host")


def build_exe():
    src = os.path.join(ROOT, "tests", "cpp", "test_host.cpp")
    libdir = os.path.dirname(q.library_path())
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    cmd = ["g++", "-std=c++14", "-O1", "-o", EXE, src, "-L" + libdir, "-lgo1mpc", "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    return EXE


def test_host_header_compiles_headless():
    q.load_library()
    build_exe()
    hdr = open(os.path.join(ROOT, "quadrupedal_loco_b200", "host", "go1mpc.hpp")).read()
    for banned in ("#include <ros", "Eigen/", "armadillo", "boost/", "fusion.h"):
        assert banned not in hdr


@pytest.mark.gpu
def test_host_classes_reproduce_oracle(oracle):
    exe = build_exe()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout)
    assert d["qp_ok"] == 1 and abs(d["qp_cost"] - 12) < 1e-12
    np.testing.assert_allclose(d["qp_x"], [1, 2], atol=1e-12)
    # body MPC, closed loop
    nh = 10
    cfg = oracle.body_cfg(nh)
    k = np.arange(nh)
    refs = np.zeros((1, 9, nh))
    refs[0, 0] = 0.30 + 0.001 * k; refs[0, 1] = 0.12; refs[0, 2] = 0.21 - 0.002 * k; refs[0, 3] = -0.19
    refs[0, 4] = 0.28; refs[0, 5] = -0.127; refs[0, 6] = 0.31; refs[0, 7] = 0.126; refs[0, 8] = 0.3 * np.sin(0.7 * k)
    from quadrupedal_loco_b200 import synth
    tx = synth.default_tx()[None, :]
    theta = np.array([[0.05, -0.4, -0.08, 0.6]]); x = np.zeros((1, 2 * nh)); o14 = np.zeros((1, 14))
    got14 = np.array(d["body_out14"]).reshape(30, 14); gotth = np.array(d["body_theta"]).reshape(30, 4)
    for t in range(30):
        oracle.body_step_batch(cfg, np.array([200 + t], np.int32), tx, theta, np.zeros((1, 4)), refs, o14, x)
        assert np.abs(got14[t] - o14[0]).max() < 1e-9 * max(1, np.abs(o14).max()), t
        assert np.abs(gotth[t] - theta[0]).max() < 1e-9, t
    # step timing replay against the reference's own golden outputs
    from tests.test_oracle_vs_ref import load
    g = load("step_ref.npz")
    got = np.array(d["step_out38"]).reshape(120, 38)
    ref = g["replay_out"][1:121]
    assert np.abs(got - ref).max() / max(1, np.abs(ref).max()) < 1e-9
    assert np.array_equal(got[:, 27], ref[:, 27]) and np.array_equal(got[:, 34], ref[:, 34])
    assert list(d["right_support"]) == list(g["replay_right_support"][1:121])
    foot = np.array(d["foot_out18"]).reshape(120, 18)
    assert np.abs(foot[:, :6] - g["replay_foot"][1:121, :6]).max() < 2e-6     # closed loop: see test_foot_trajectory_replay...
    # the two nodes in lock step: slow messages vs the unmodified NLPRTControlClass up to the stop (tick 260), the stopped tail and
    # the 100 Hz node's messages vs the oracle chains fed the same way
    from tests.test_oracle_vs_ref import OracleNlpNode, OracleRtNode
    gs = load("rt_node_ref.npz")
    slow = np.array(d["node_slow"]).reshape(400, 100); fast = np.array(d["node_fast"]).reshape(d["node_nfast"], 100)
    rel = lambda a, b: (np.abs(a - b) / np.maximum(1.0, np.abs(b))).max()
    assert rel(slow[:259], gs["msgs"][1:260]) < 1e-9 and np.array_equal(slow[:259, [27, 97, 99]], gs["msgs"][1:260][:, [27, 97, 99]])
    ns, nf = OracleNlpNode(oracle), OracleRtNode(oracle, 4)
    want_slow = np.zeros((400, 100)); want_fast = []; msg = np.zeros(100); count = 0; t_ms = 0
    while count < 400 or t_ms % 25:
        if t_ms % 25 == 0:
            count += 1
            if count == 260:
                ns.stop()
            msg = ns.tick(count, 1, np.zeros(3), np.zeros(3)); want_slow[count - 1] = msg
        if t_ms % 10 == 0 and len(want_fast) < 1001:
            want_fast.append(nf.tick(msg))
        t_ms += 5
    want_fast = np.array(want_fast)
    assert rel(slow, want_slow) < 1e-9 and slow[300:, [8, 11]].max() == 0.0            # lift heights zeroed after the stop
    assert fast.shape == want_fast.shape and rel(fast, want_fast) < 1e-9
    assert np.array_equal(fast[:, [27, 63, 98, 99]], want_fast[:, [27, 63, 98, 99]])
    assert (np.abs(fast[:, 72:86]).sum(axis=1) > 0).sum() > 800                          # the body MPC really ran
    # Dynamiccclass chain: closed-form split -> force QP -> torque map, twice (pace, then trot from the first grf_opt)
    from tests.test_grf import oracle_grf, oracle_tau
    hom = np.array([0.1881, -0.1268, 0, 0.1881, 0.1268, 0, -0.1881, -0.1268, 0, -0.1881, 0.1268, 0])
    prev = np.zeros(12)
    for c in range(2):
        dd = dict(mode=np.array([101 + c], np.int32), rs=np.array([c], np.int32), base=np.array([[0.01, -0.02, 0.31]]),
                  legs=(hom + 0.01 * np.sin(1.3 * np.arange(12)))[None, :], FT=np.array([[5, -7, 117.6, 1, -2, 0.5]]),
                  F6=np.array([[3.0, -2, 60, -4, 5, 55]]), rf=np.array([[0.01, -0.127, 0]]), lf=np.array([[-0.01, 0.127, 0]]), prev=prev[None, :].copy())
        o = oracle_grf(oracle, dd)
        np.testing.assert_allclose(np.array(d["grf_guess"])[12 * c:12 * c + 12], o["Fg"][0], rtol=1e-12, atol=1e-12)
        assert d["grf_ok"][c] == o["ok"][0] == 1
        np.testing.assert_allclose(np.array(d["grf_opt"])[12 * c:12 * c + 12], o["grf"][0], rtol=1e-9, atol=1e-9)
        prev = o["grf"][0]
        kk = np.arange(9)
        td = dict(jac=np.stack([0.1 * np.cos(0.9 * kk + l) for l in range(4)])[None], swing=np.array([[(l + c) % 2 == 0 for l in range(4)]], np.int32),
                  p_des=np.tile(0.1 * np.arange(3), (1, 4, 1)), p_est=np.tile(0.1 * np.arange(3) + 0.01, (1, 4, 1)),
                  pv_des=np.full((1, 4, 3), 0.2), pv_est=np.tile(0.15 - 0.1 * np.arange(3), (1, 4, 1)),
                  F=np.array(d["grf_guess"])[12 * c:12 * c + 12].reshape(1, 4, 3))      # F_leg_ref stays the closed-form split
        np.testing.assert_allclose(np.array(d["grf_tau"])[12 * c:12 * c + 12].reshape(4, 3), oracle_tau(oracle, td)[0], rtol=0, atol=1e-12)
    # kinematics
    pos, J = oracle.leg_fk(np.array([[0.1, 0.8, -1.5]]), [1], np.array([[0, 0, 0.31]]), np.array([[0.05, -0.04, 0.1]]))
    np.testing.assert_allclose(d["fk_pos"], pos[0], atol=1e-13)
    np.testing.assert_allclose(np.array(d["fk_J"]).reshape(3, 3).T.ravel(), J[0], atol=1e-13)   # Mat is column-major
    qs, _, it = oracle.leg_ik(pos, np.array([[0.0, 0.87, -1.5]]), [1], np.array([[0, 0, 0.31]]), np.array([[0.05, -0.04, 0.1]]))
    np.testing.assert_allclose(d["ik_q"], qs[0], atol=1e-9)
    assert d["ik_updates"] == it[0]
