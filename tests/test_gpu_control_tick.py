"""The end-to-end entry bench.py's e2e leg times: go1mpc_control_tick_host_async (host inputs up, planner tick + body tick on
the device-resident records, 12-double result rows down) against the CPU oracle DIRECTLY -- orc_step_timing_tick (NLPClass::
step_timing_opti_loop) and orc_body_theta_mpc (PRMPCClass::body_theta_mpc) on the same inputs: the compact rows, the full
planner outputs and both diagnostics.  Also: the 10-row sensor upload (step_in_rows = 10) equals the 20-row one, including
after a change of batch size on the same stream."""
import ctypes

import numpy as np
import pytest

import quadrupedal_loco_b200 as q
from quadrupedal_loco_b200 import synth
from tests.test_gpu_body import assert_body_parity
from tests.test_gpu_step import assert_step_parity

pytestmark = pytest.mark.gpu


def control_tick(mpc, nh, d, tick, st, sin, rows=0, stream=None):
    """One control tick through the C ABI with pinned host buffers; returns compact rows, out38, step diag, body out, body diag,
    planner state after the tick."""
    import torch
    dev = torch.device("cuda", 0)
    B = len(tick)
    rec = q.pack_body_inputs(nh, d["tick"], d["tx"], d["theta"], d["bstate"], d["x_warm"], d["refs"])
    tx, xw, trec = q.split_body_record(nh, rec)
    os_, ds = q.body_out_stride(nh), q.body_diag_stride(nh)
    tx_d = torch.from_numpy(tx).to(dev)
    bout = np.zeros((B, os_)); bout[:, 18:18 + 2 * nh] = xw
    bout_d = torch.from_numpy(bout).to(dev)
    st_d = torch.from_numpy(np.ascontiguousarray(st.T)).to(dev)
    comp_d = torch.zeros(B, q.COMPACT_DOUBLES, dtype=torch.float64, device=dev)
    tick_h = torch.from_numpy(np.ascontiguousarray(tick, np.int32)).pin_memory()
    sin_soa = np.ascontiguousarray(sin.T)
    sin_h = torch.from_numpy(sin_soa[:rows] if rows else sin_soa).contiguous().pin_memory()
    trec_h = torch.from_numpy(trec).pin_memory()
    comp_h = torch.zeros(B, q.COMPACT_DOUBLES, dtype=torch.float64).pin_memory()
    o38_h = torch.zeros(q.STEP_OUT, B, dtype=torch.float64).pin_memory()
    sdg_h = torch.full((q.STEP_DIAG, B), -7, dtype=torch.int32).pin_memory()
    bdg_h = torch.full((B, ds), -7, dtype=torch.int32).pin_memory()
    t = q.ControlTick()
    t.n_sqp, t.nh = 3, nh
    t.tick, t.step_in, t.body_tick_in = tick_h.data_ptr(), sin_h.data_ptr(), trec_h.data_ptr()
    t.step_state_src_d = None; t.step_state_d = st_d.data_ptr()
    t.tx_d, t.body_out_d, t.compact_d = tx_d.data_ptr(), bout_d.data_ptr(), comp_d.data_ptr()
    t.compact, t.out38, t.step_diag, t.body_diag = comp_h.data_ptr(), o38_h.data_ptr(), sdg_h.data_ptr(), bdg_h.data_ptr()
    t.step_in_rows = rows
    torch.cuda.synchronize()
    rc = mpc.lib.go1mpc_control_tick_host_async(mpc.h, B, ctypes.byref(t), stream)
    assert rc == 0, mpc.last_error() if hasattr(mpc, "last_error") else rc
    torch.cuda.synchronize()
    return (comp_h.numpy().copy(), o38_h.numpy().T.copy(), sdg_h.numpy().T.copy(), bout_d.cpu().numpy(), bdg_h.numpy().copy(),
            st_d.cpu().numpy().T.copy())


def workload(mpc, B, nh, seed):
    d = synth.body_mpc_inputs(B, nh, seed=seed, scale=2.0, theta_clip=0.16)
    tick, st, sin = synth.step_timing_inputs(B, mpc.step_default_state(), seed=seed, amp=2.0, push_x=0.4, push_y=0.75, p_hi=16)
    return d, tick, st, sin


@pytest.mark.parametrize("B", [1, 777, 5000])
def test_control_tick_vs_oracle(mpc, oracle, B):
    nh = 10
    d, tick, st, sin = workload(mpc, B, nh, synth.SEED_CFG3 + B)
    comp, o38, sdg, bout, bdg, st_after = control_tick(mpc, nh, d, tick, st, sin)
    # planner tick
    scfg = oracle.step_cfg(3)
    os_ = st.copy()
    oo, od = oracle.step_tick_batch(scfg, tick, os_, sin)
    ok = assert_step_parity(o38, st_after, sdg, oo, os_, od, f"control tick planner B={B}")
    # body tick
    cfg = oracle.body_cfg(nh)
    theta = d["theta"].copy(); x = d["x_warm"].copy(); o14 = np.zeros((B, 14))
    r = oracle.body_step_batch(cfg, d["tick"], d["tx"], theta, d["bstate"], d["refs"], o14, x)
    r.update(theta=theta, x=x, out14=o14)
    assert_body_parity(bout, bdg, r, nh, f"control tick body B={B}")
    # compact rows: [0,3) CoM | [3,5) roll pitch | [5,7) torques | [7,9) next footstep | [9] period | [10] planner status | [11] body status
    rel = lambda a, b: np.abs(a - b) / np.maximum(1.0, np.abs(b))
    assert rel(comp[ok, 0:3], oo[ok, 0:3]).max() < 1e-9
    assert rel(comp[:, 3:7], o14[:, 0:4]).max() < 1e-9
    assert rel(comp[ok][:, [7, 8, 9]], oo[ok][:, [29, 31, 35]]).max() < 1e-9
    np.testing.assert_array_equal(comp[:, 11], r["status"])
    last = np.array([sdg[b, 5 + 11 * 2] for b in range(B)])            # status of the third (last) SQP solve
    np.testing.assert_array_equal(comp[:, 10], last)


def test_control_tick_sensor_rows_only(mpc):
    """step_in_rows = 10 (estimated CoM state + foot locations; the rows the planner reads without external heights on flat
    ground) gives the same results as the 20-row upload -- also when the batch size changes on the same stream, where stale
    sensor rows of the larger batch would otherwise alias the zero rows of the smaller one."""
    import torch
    nh = 10
    s = torch.cuda.Stream()
    for B in (4000, 1500, 4000):
        d, tick, st, sin = workload(mpc, B, nh, 99 + B)
        sin[:, 10:13] = 0.309458                                       # what synth leaves there: unused without ext_height
        full = control_tick(mpc, nh, d, tick, st, sin, rows=0, stream=s.cuda_stream)
        part = control_tick(mpc, nh, d, tick, st, sin, rows=10, stream=s.cuda_stream)
        for a, b_ in zip(full, part):
            np.testing.assert_array_equal(a, b_)
