"""Both MPC nodes chained ON THE DEVICE, in lock step, with no host in between ("the closed 40 Hz -> 100 Hz chain
fully on device"): go1mpc_nlp_node_tick_batch writes the /MPC/Gait buffer every 25 ms, go1mpc_rt_node_tick_batch reads it every
10 ms.  Golden: the same schedule run on the UNMODIFIED NLPRTControlClass and PRMPCClass (tests/golden/rt_node_ref.npz, cfg1:
719 slow ticks, 1798 fast ticks, nh = 4) -- every /rtMPC/traj message to 1e-8 relative (the slow node's 1e-12-level
differences pass through the fast node's cubic fits), integer slots exact."""
import numpy as np
import pytest

import quadrupedal_loco_b200 as q
from tests.test_oracle_vs_ref import load

pytestmark = pytest.mark.gpu


def test_device_chain_lockstep_vs_unmodified_classes(mpc):
    import torch
    g = load("rt_node_ref.npz")
    nh = int(g["nh"][0])
    want, mo = g["out"], g["msg_of_fast"]
    n_slow = len(g["msgs"]) - 1
    dev = torch.device("cuda", 0)
    R = 4
    f64 = dict(dtype=torch.float64, device=dev)
    nlp_st = torch.from_numpy(np.repeat(mpc.nlp_node_default_state()[:, None], R, axis=1).copy()).to(dev)
    rt_st = torch.from_numpy(np.repeat(mpc.rt_node_default_state(nh)[:, None], R, axis=1).copy()).to(dev)
    body_in = torch.zeros(R, q.body_in_stride(nh), **f64); body_out = torch.zeros(R, q.body_out_stride(nh), **f64)
    gait = torch.zeros(100, R, **f64)                                   # the /MPC/Gait buffer both nodes share
    outs = torch.zeros(len(want), 100, R, **f64)
    wd = (torch.arange(n_slow + 1, dtype=torch.int32)[:, None] * torch.ones(1, R, dtype=torch.int32)).contiguous().to(dev)   # [tick][robot]
    torch.cuda.synchronize()
    count, k, t_ms = 0, 0, 0
    while count < n_slow or t_ms % 25:
        if t_ms % 25 == 0:
            count += 1
            mpc.nlp_node_tick(R, nlp_st, wd[count], gait)       # everything on the handle's stream: the nodes are ordered
        if t_ms % 10 == 0:
            assert mo[k] == count
            mpc.rt_node_tick(nh, R, rt_st, gait, body_in, body_out, outs[k])
            k += 1
        t_ms += 5
    mpc.synchronize(); torch.cuda.synchronize()
    assert k == len(want)
    got = outs.cpu().numpy().transpose(0, 2, 1)
    ints = [27, 63, 98, 99]
    for r in range(R):
        np.testing.assert_array_equal(got[:, r][:, ints], want[:, ints])
        err = np.abs(got[:, r] - want) / np.maximum(1.0, np.abs(want))
        assert np.isfinite(got).all() and err.max() < 1e-8, f"robot {r}: {err.max():.3e} at {np.unravel_index(np.argmax(err), err.shape)}"
    assert (np.abs(want[:, 72:86]).sum(axis=1) > 0).sum() > 1500
