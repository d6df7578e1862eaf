"""The 100 Hz node on the device (go1mpc_rt_node_tick_batch, csrc/rt_chain.cu): cfg1 lock-step replay.

Golden: tests/golden/rt_node_ref.npz -- the UNMODIFIED NLPRTControlClass::WalkingReactStepping (40 Hz, 40 squat ticks + the
671-tick walk) publishing /MPC/Gait, and the UNMODIFIED PRMPCClass behind gait_fast.cpp's glue (RT/src/gait_fast.cpp:505-746)
consuming it every 10 ms, 1798 fast ticks at the reference's horizon nh = 4.  The device chain gets the same messages and must
reproduce every outgoing /rtMPC/traj message to 1e-9 (relative to max(1, |value|)) with the integer slots (27 step index,
98, 99 counters) exact.  At nh = 10, which the reference cannot compile, the checker is the oracle chain
(tests/test_oracle_vs_ref.py::OracleRtNode, pinned bit for bit against the same golden at nh = 4), on robots whose message
streams are shifted against each other."""
import numpy as np
import pytest

import quadrupedal_loco_b200 as q
from tests.test_oracle_vs_ref import OracleRtNode, load

pytestmark = pytest.mark.gpu


def run_device(mpc, nh, msgs_per_tick):
    """msgs_per_tick: [T, R, 100] the message each of R robots sees at fast tick t.  Returns [T, R, 100]."""
    import torch
    dev = torch.device("cuda", 0)
    T, R, _ = msgs_per_tick.shape
    S = mpc.rt_node_state_doubles(nh)
    st = torch.from_numpy(np.repeat(mpc.rt_node_default_state(nh)[:, None], R, axis=1).copy()).to(dev)
    body_in = torch.zeros(R, q.body_in_stride(nh), dtype=torch.float64, device=dev)
    body_out = torch.zeros(R, q.body_out_stride(nh), dtype=torch.float64, device=dev)
    msg_all = torch.from_numpy(np.ascontiguousarray(msgs_per_tick.transpose(0, 2, 1))).to(dev)     # [T, 100, R]
    res = torch.zeros(T, 100, R, dtype=torch.float64, device=dev)
    assert st.shape[0] == S
    torch.cuda.synchronize()
    for t in range(T):
        mpc.rt_node_tick(nh, R, st, msg_all[t], body_in, body_out, res[t])      # every tick on the handle's stream
    mpc.synchronize()
    torch.cuda.synchronize()
    return res.cpu().numpy().transpose(0, 2, 1)


def assert_messages(got, want, what):
    ints = [27, 63, 98, 99]                                   # step index (and its copy in the interpolated block), counters
    np.testing.assert_array_equal(got[..., ints], want[..., ints], err_msg=what + " (integer slots)")
    err = np.abs(got - want) / np.maximum(1.0, np.abs(want))
    k = np.unravel_index(np.argmax(err), err.shape)
    assert np.isfinite(got).all() and err.max() < 1e-9, f"{what}: max rel err {err.max():.3e} at tick/robot/slot {k}"


def test_rt_node_lockstep_replay_vs_unmodified_classes(mpc):
    g = load("rt_node_ref.npz")
    nh = int(g["nh"][0])
    msgs, want, mo = g["msgs"], g["out"], g["msg_of_fast"]
    R = 5
    per_tick = np.repeat(msgs[mo][:, None, :], R, axis=1)
    got = run_device(mpc, nh, per_tick)
    assert (np.abs(want[:, 72:86]).sum(axis=1) > 0).sum() > 1500            # the body MPC really ran
    for r in range(R):
        assert_messages(got[:, r], want, f"nh={nh} robot {r}")


def test_rt_node_nh10_vs_oracle_chain(mpc, oracle):
    g = load("rt_node_ref.npz")
    msgs, mo = g["msgs"], g["msg_of_fast"]
    nh, R, T = 10, 3, 900
    shifts = [0, 7, 23]                                       # robots out of phase with each other
    per_tick = np.zeros((T, R, 100))
    for r, sft in enumerate(shifts):
        idx = np.clip(np.arange(T) + 300 - sft, 0, len(mo) - 1)
        per_tick[:, r] = msgs[mo[idx]]
    got = run_device(mpc, nh, per_tick)
    for r in range(R):
        node = OracleRtNode(oracle, nh)
        want = np.array([node.tick(per_tick[t, r]) for t in range(T)])
        assert (np.abs(want[:, 72:86]).sum(axis=1) > 0).sum() > 500
        assert_messages(got[:, r], want, f"nh=10 robot {r}")


def test_rt_node_wire_format_topics(mpc, oracle):
    """go1mpc_rt_node_tick_msgs_batch: /control2rtmpc/state (25 slots) in, /rt2nrt/state out (gait_fast.cpp:92-110, 519-527).
    It must equal the entry that takes the control flag and bodyangle_state separately, bit for bit (messages and state), feed
    the body MPC the measured angles of slots 10, 11, 13, 14, and match the oracle chain given the same flag and angles."""
    import torch
    g = load("rt_node_ref.npz")
    msgs, mo = g["msgs"], g["msg_of_fast"]
    nh, R, T = 4, 3, 400
    dev = torch.device("cuda", 0)
    rng = np.random.Generator(np.random.Philox(12))
    ctl = np.zeros((T, 25, R))
    ctl[:, 0] = 1.0; ctl[100:110, 0, 1] = 0.0                       # robot 1: control off for ten ticks
    ctl[:, 1:] = rng.uniform(-0.02, 0.02, (T, 24, R))
    per_tick = np.repeat(msgs[mo[300:300 + T]][:, :, None], R, axis=2)        # [T, 100, R]

    def fresh():
        return (torch.from_numpy(np.repeat(mpc.rt_node_default_state(nh)[:, None], R, axis=1).copy()).to(dev),
                torch.zeros(R, q.body_in_stride(nh), dtype=torch.float64, device=dev),
                torch.zeros(R, q.body_out_stride(nh), dtype=torch.float64, device=dev))
    m_d = torch.from_numpy(per_tick).to(dev); c_d = torch.from_numpy(ctl).to(dev)
    outA = torch.zeros(T, 100, R, dtype=torch.float64, device=dev); outB = torch.zeros_like(outA)
    r2n = torch.zeros(T, 25, R, dtype=torch.float64, device=dev)
    stA, biA, boA = fresh(); stB, biB, boB = fresh()
    flag = torch.from_numpy((ctl[:, 0] > 0).astype(np.int32)).to(dev)
    bs = torch.from_numpy(np.ascontiguousarray(ctl[:, [10, 11, 13, 14]])).to(dev)
    torch.cuda.synchronize()
    for t in range(T):
        mpc.rt_node_tick_msgs(nh, R, stA, m_d[t], c_d[t], biA, boA, outA[t], r2n[t])
        mpc.rt_node_tick(nh, R, stB, m_d[t], biB, boB, outB[t], ctrl_d=flag[t], bodyangle_state_d=bs[t])
    mpc.synchronize(); torch.cuda.synchronize()
    a, b_ = outA.cpu().numpy(), outB.cpu().numpy()
    np.testing.assert_array_equal(a, b_)
    np.testing.assert_array_equal(stA.cpu().numpy(), stB.cpu().numpy())
    np.testing.assert_array_equal(biA.cpu().numpy()[:, 32:36], ctl[T - 1][[10, 11, 13, 14]].T)      # what the body MPC was fed
    r = r2n.cpu().numpy()
    np.testing.assert_array_equal(r[:, 1:], ctl[:, 1:])
    np.testing.assert_array_equal(r[-1, 0], stA.cpu().numpy()[3])                                  # slot 0 = t_int
    assert (np.diff(r[:, 0, 0]) >= 0).all() and r[-1, 0, 0] > 0 and r[-1, 0, 1] < r[-1, 0, 0]         # robot 1 lost ten ticks
    for rb in (0, 1):
        node = OracleRtNode(oracle, nh)
        want = np.array([node.tick(per_tick[t, :, rb], ctrl=int(ctl[t, 0, rb] > 0), bs=np.ascontiguousarray(ctl[t, [10, 11, 13, 14], rb]))
                         for t in range(T)])
        assert_messages(a[:, :, rb], want, f"msgs entry robot {rb}")
