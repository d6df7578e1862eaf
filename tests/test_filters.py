"""Signal filters of the servo loop: butterworthLPF and ButterworthFilter::ForceFilter
(GO1/src/Filter/butterworthLPF.cpp:82-121, butterworth_filter.cpp:37-69).

CPU: the oracle (oracle/filters.c) and the library's host-side coefficients against the UNMODIFIED classes, bit for bit
(tests/golden/filter_ref.npz + live where oracle/_ref is present).  GPU: go1mpc_lpf_batch / go1mpc_force_filter_batch for a
batch of robots x channels against the oracle, bit for bit (streaming kernels, -fmad=false, the reference's summation order),
with the channels picked out of a [100][B] message buffer as go1_servo picks them (servo.cpp:898-931)."""
import ctypes

import numpy as np
import pytest

import quadrupedal_loco_b200 as q
from tests.oracle_lib import P, ref_path
from tests.test_oracle_vs_ref import load


class Coef(ctypes.Structure):
    _fields_ = [(n, ctypes.c_double) for n in ("b0", "b1", "b2", "a1", "a2", "a")]


def oracle_lpf(oracle, fs, fc, x):
    lib = oracle.lib
    lib.orc_lpf_filter.restype = ctypes.c_double
    lib.orc_lpf_filter.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double]
    lib.orc_lpf_init.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_void_p]
    c = Coef(); lib.orc_lpf_init(float(fs), float(fc), ctypes.byref(c))
    st = np.zeros(5)
    return np.array([lib.orc_lpf_filter(ctypes.byref(c), P(st), float(v)) for v in x]), c


def oracle_force(oracle, x):
    lib = oracle.lib
    lib.orc_force_filter.restype = ctypes.c_double
    lib.orc_force_filter.argtypes = [ctypes.c_void_p, ctypes.c_double]
    st = np.zeros(6)
    return np.array([lib.orc_force_filter(P(st), float(v)) for v in x])


def test_oracle_and_host_coefficients_bit_exact_vs_reference_golden(oracle):
    g = load("filter_ref.npz")
    x, fsfc = g["x"], g["fsfc"]
    for c in range(x.shape[1]):
        y, k = oracle_lpf(oracle, fsfc[c, 0], fsfc[c, 1], x[:, c])
        np.testing.assert_array_equal(y, g["lpf"][:, c])
        np.testing.assert_array_equal([k.b0, k.b1, k.b2, k.a1, k.a2, k.a], g["coefs"][c, :6])
        np.testing.assert_array_equal(q.lpf_coefficients(fsfc[c, 0], fsfc[c, 1]), g["coefs"][c, :6])    # library, host side
        np.testing.assert_array_equal(oracle_force(oracle, 50 * x[:, c]), g["force"][:, c])
    assert np.abs(g["lpf"][50:, 0] - x[50:, 0]).max() > 0.05          # the filter does something


def test_oracle_filters_live_vs_reference(oracle):
    path = ref_path("libref_filter.so")
    if not path:
        pytest.skip("oracle/_ref absent")
    rf = ctypes.CDLL(path)
    rf.ref_lpf_new.restype = ctypes.c_void_p; rf.ref_lpf_new.argtypes = [ctypes.c_double] * 2
    rf.ref_lpf_filter.restype = ctypes.c_double; rf.ref_lpf_filter.argtypes = [ctypes.c_void_p, ctypes.c_double]
    rng = np.random.Generator(np.random.Philox(5))
    for fs, fc in ((1000.0, 3.0), (1000.0, 10.0), (500.0, 40.0), (1000.0, 499.0)):
        x = rng.standard_normal(300).cumsum() * 0.01
        h = ctypes.c_void_p(rf.ref_lpf_new(fs, fc))
        want = np.array([rf.ref_lpf_filter(h, float(v)) for v in x])
        np.testing.assert_array_equal(oracle_lpf(oracle, fs, fc, x)[0], want)


@pytest.mark.gpu
def test_filters_device_bit_exact(mpc, oracle):
    import torch
    dev = torch.device("cuda", 0)
    rng = np.random.Generator(np.random.Philox(8))
    B, T = 257, 60
    # go1_servo's channels: message slots 0-2 (CoM), 9-11 / 6-8 (feet) at 3 Hz, 39-41 (CoM acceleration), 73-75 at 10 Hz
    rows = np.array([0, 1, 2, 9, 10, 11, 6, 7, 8, 39, 40, 41, 73, 74, 75], np.int32)
    fc = np.array([3.0] * 9 + [10.0] * 6)
    C = len(rows)
    coef = np.stack([q.lpf_coefficients(1000.0, f) for f in fc])
    msgs = rng.standard_normal((T, 100, B)).cumsum(axis=0) * 0.01
    st = torch.zeros(5, C, B, dtype=torch.float64, device=dev); out = torch.zeros(T, C, B, dtype=torch.float64, device=dev)
    m_d = torch.from_numpy(msgs).to(dev); rows_d = torch.from_numpy(rows).to(dev)
    fst = torch.zeros(6, C, B, dtype=torch.float64, device=dev); fout = torch.zeros(T, C, B, dtype=torch.float64, device=dev)
    fin = torch.from_numpy(np.ascontiguousarray(msgs[:, rows, :] * 40)).to(dev)
    torch.cuda.synchronize()
    for t in range(T):
        mpc.lpf_batch(B, C, coef, m_d[t], st, out[t], in_rows_d=rows_d)
        mpc.force_filter_batch(B, C, fin[t], fst, fout[t])
    mpc.synchronize()
    got = out.cpu().numpy(); gotf = fout.cpu().numpy()
    for c in range(C):
        for b in (0, 1, 100, B - 1):
            np.testing.assert_array_equal(got[:, c, b], oracle_lpf(oracle, 1000.0, fc[c], msgs[:, rows[c], b])[0])
            np.testing.assert_array_equal(gotf[:, c, b], oracle_force(oracle, msgs[:, rows[c], b] * 40))
    assert st.cpu().numpy()[0].max() == 3.0            # the call counter stops at 3
    mpc.lpf_batch(0, C, coef, m_d[0], st, out[0])      # empty batch: no-op
