"""quadrupedal_loco_b200 -- B200-native batched Go1 gait-planning MPC (host-side Python binding).

The product is libgo1mpc.so (CUDA, sm_100a) behind the C ABI in include/go1mpc.h; this
package is the thin ctypes binding tests and bench.py use.  PyTorch is only used by
callers for device memory and streams.  There is no CPU fallback: if the library is
missing or no CUDA device is usable, calls raise.
"""
from .binding import (  # noqa: F401
    Go1Mpc, Go1MpcError, load_library, library_path, body_in_stride, body_out_stride,
    body_diag_stride, pack_body_inputs, EXPORTED_SYMBOLS, QP_ITERS, BODY_DIAG_ACTIVE,
    STEP_STATE, STEP_IN, STEP_OUT, STEP_DIAG, BODY_TICK_OUT, body_tick_in_stride, split_body_record,
    ControlTick, COMPACT_DOUBLES, lpf_coefficients,
)
from ._build import build  # noqa: F401

__all__ = [
    "Go1Mpc", "Go1MpcError", "load_library", "library_path", "build", "body_in_stride",
    "body_out_stride", "body_diag_stride", "pack_body_inputs", "EXPORTED_SYMBOLS", "QP_ITERS",
    "BODY_DIAG_ACTIVE", "STEP_STATE", "STEP_IN", "STEP_OUT", "STEP_DIAG", "BODY_TICK_OUT", "body_tick_in_stride",
    "split_body_record", "ControlTick", "COMPACT_DOUBLES", "lpf_coefficients",
]
