"""The 40 Hz planner node on the device (go1mpc_nlp_node_tick_batch, csrc/nlp_chain.cu).

Golden: the UNMODIFIED NLPRTControlClass::WalkingReactStepping (oracle/_ref/libref_nlp.so) -- cfg1's message sequence
(tests/golden/rt_node_ref.npz: 40 squat ticks, the 671-tick walk, 8 ticks beyond its end) and scripted stop / restart / idle
sequences with noisy foot-location feedback (tests/golden/nlp_node_ref.npz).  Every robot of the batch runs a different script,
so one launch mixes squat, walking, stopped, frozen and finished robots.  Contract: every message slot to 1e-9 (relative to
max(1, |value|)), the integer slots (27 step index, 97 stop flag, 99 support) exact.  The device chain's state layout is the
oracle's (oracle/nlp_node.c, pinned bit for bit against the same goldens), compared at the end."""
import numpy as np
import pytest

from tests.golden.make_golden import nlp_node_feedback, nlp_node_scripts
from tests.test_oracle_vs_ref import OracleNlpNode, load, run_nlp_script

pytestmark = pytest.mark.gpu
CMD = {"stop": 1, "start": 2}


def run_device(mpc, scripts, T):
    """scripts: list of (events, idle, rf [T+1,3], lf [T+1,3]); returns messages [T+1, R, 100] and the final state [S, R]."""
    import torch
    dev = torch.device("cuda", 0)
    R = len(scripts)
    st = torch.from_numpy(np.repeat(mpc.nlp_node_default_state()[:, None], R, axis=1).copy()).to(dev)
    out = torch.zeros(T + 1, 100, R, dtype=torch.float64, device=dev)
    wd = torch.zeros(T + 1, R, dtype=torch.int32); start = torch.ones(T + 1, R, dtype=torch.int32); cmd = torch.zeros(T + 1, R, dtype=torch.int32)
    rf = torch.zeros(T + 1, 3, R, dtype=torch.float64); lf = torch.zeros(T + 1, 3, R, dtype=torch.float64)
    for r, (events, idle, rfr, lfr) in enumerate(scripts):
        n = min(T + 1, len(rfr))
        wd[:, r] = torch.arange(T + 1, dtype=torch.int32)
        for t, ev in events.items():
            cmd[t, r] = CMD[ev]
        for t in idle:
            start[t, r] = 0
        rf[:n, :, r] = torch.from_numpy(rfr[:n]); lf[:n, :, r] = torch.from_numpy(lfr[:n])
    wd, start, cmd, rf, lf = (x.to(dev) for x in (wd, start, cmd, rf, lf))
    torch.cuda.synchronize()
    for t in range(1, T + 1):
        mpc.nlp_node_tick(R, st, wd[t], out[t], start_d=start[t], cmd_d=cmd[t], rfoot_fb_d=rf[t], lfoot_fb_d=lf[t])
    mpc.synchronize()
    torch.cuda.synchronize()
    return out.cpu().numpy().transpose(0, 2, 1), st.cpu().numpy()


def assert_messages(got, want, what):
    ints = [27, 97, 99]
    np.testing.assert_array_equal(got[..., ints], want[..., ints], err_msg=what + " (integer slots)")
    err = np.abs(got - want) / np.maximum(1.0, np.abs(want))
    k = np.unravel_index(np.argmax(err), err.shape)
    assert np.isfinite(got).all() and err.max() < 1e-9, f"{what}: max rel err {err.max():.3e} at tick/slot {k}"


def test_nlp_node_vs_unmodified_class(mpc):
    g = load("rt_node_ref.npz"); g2 = load("nlp_node_ref.npz")
    T = len(g["msgs"]) - 1
    z = np.zeros((T + 1, 3))
    scripts = [({}, (), z, z)]
    wants = [g["msgs"]]
    for name, Ts, events, idle, seed in nlp_node_scripts():
        rf, lf = nlp_node_feedback(seed, Ts)
        scripts.append((events, idle, rf, lf)); wants.append(g2[name])
    got, _ = run_device(mpc, scripts, T)
    assert mpc.nlp_walkdtime_max() == 672
    for r, want in enumerate(wants):
        n = min(T + 1, len(want))
        assert_messages(got[1:n, r], want[1:n], f"robot {r}")
    assert wants[3][260:, [8, 11]].max() == 0.0 and got[260:420, 3][:, [8, 11]].max() == 0.0      # the stop zeroed the lift heights


def test_nlp_node_state_matches_oracle_chain(mpc, oracle):
    """Batch of out-of-phase robots (different start ticks, stops, feedback) against the oracle chain, state included."""
    rng = np.random.Generator(np.random.Philox(31))
    T, R = 300, 6
    scripts = []
    for r in range(R):
        rf = np.zeros((T + 1, 3)); lf = np.zeros((T + 1, 3))
        rf[:, :2] = rng.uniform(-0.015, 0.015, (T + 1, 2)); lf[:, :2] = rng.uniform(-0.015, 0.015, (T + 1, 2))
        events = {} if r % 2 == 0 else {int(rng.integers(80, 250)): "stop"}
        idle = tuple(range(1, 1 + 5 * r))
        scripts.append((events, idle, rf, lf))
    got, st = run_device(mpc, scripts, T)
    for r, (events, idle, rf, lf) in enumerate(scripts):
        node = OracleNlpNode(oracle)
        want = run_nlp_script(node, T, events, idle, rf, lf)
        assert_messages(got[1:, r], want[1:], f"robot {r}")
        n = len(node.node)
        ring = slice(234, 362)
        err = np.abs(st[:n, r] - node.node) / np.maximum(1.0, np.abs(node.node))
        err[ring] = 0.0          # ring slots older than the window differ by construction (device zeroes gaps first)
        assert err.max() < 1e-9, f"robot {r}: state row {int(np.argmax(err))} differs by {err.max():.3e}"
        live = [(int(node.node[362]) - k) & 63 for k in range(0, 50)]
        for row in (0, 1):
            a = st[234 + 64 * row + np.array(live), r]; b_ = node.node[234 + 64 * row + np.array(live)]
            assert np.abs(a - b_).max() < 1e-9
