"""The configurations bench.py times, compared with the CPU oracle DIRECTLY: the exact inputs of
bench.py's slot 0 -- cfg2 (4096 robots, gains 0) and cfg3 (65536 robots, feedback-gain presets, 2x perturbation) -- through the
same entries on the default dispatch (three-launch body path, three-launch planner path), every robot against
orc_body_theta_mpc / orc_step_timing_tick: primal 1e-9, identical ordered active sets and counters, bit-exact indices."""
import types

import numpy as np
import pytest

import quadrupedal_loco_b200 as q
from tests.test_gpu_body import assert_body_parity, run_gpu
from tests.test_gpu_step import assert_step_parity, gpu_tick

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("config,batch", [("cfg2", 4096), ("cfg3", 65536)])
def test_benched_configuration_vs_oracle(oracle, config, batch):
    import bench
    a = types.SimpleNamespace(config=config, batch=batch, nh=10)
    p = bench.plan(a, 1)
    assert p["B_global"] == batch
    mpc = q.Go1Mpc(0, {"lamda": p["body_lamda"], "step": {"lamda": p["step_lamda"]}})
    try:
        scfg = oracle.step_cfg(3, lamda=p["step_lamda"])
        body, tick, st, sin = bench.make_inputs(a, 0, 1, 1, mpc.step_default_state())
        d, tick, st, sin = body[0], tick[0], st[0], sin[0]
        # body-inclination MPC tick
        out, diag = run_gpu(mpc, 10, d, device=True)
        cfg = oracle.body_cfg(10, lamda=p["body_lamda"])
        theta = d["theta"].copy(); x = d["x_warm"].copy(); o14 = np.zeros((batch, 14))
        r = oracle.body_step_batch(cfg, d["tick"], d["tx"], theta, d["bstate"], d["refs"], o14, x)
        r.update(theta=theta, x=x, out14=o14)
        assert_body_parity(out, diag, r, 10, f"{config} body")
        assert mpc.body_guard_trips() == 0
        # step-location / step-timing SQP tick
        go, gs, gd = gpu_tick(mpc, tick, st, sin, 3, device=True)
        os_ = st.copy()
        oo, od = oracle.step_tick_batch(scfg, tick, os_, sin)
        ok = assert_step_parity(go, gs, gd, oo, os_, od, f"{config} planner")
        assert ok.mean() > 0.85
    finally:
        mpc.close()
