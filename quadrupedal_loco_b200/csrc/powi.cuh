// powi.cuh -- correctly rounded small integer powers (shared by the time-polynomial fits).
#pragma once
#include <cuda_runtime.h>

namespace go1 {

// x / d, bit-identical to the IEEE division, without the detour a zero numerator costs: CUDA's inline FP64 division
// only covers operands of ordinary magnitude and sends everything else -- including the exact zeros that fill the
// triangular / sparse matrices of these solvers -- to a ~100-instruction subroutine (18 % of the planner kernel's
// instructions before this).  0 / d = +-0 for every finite non-zero or infinite d; NaN and 0 / 0 take the real division.
// (not inlined: ~60 call sites in the planner kernel, whose 230 KB of code stall on instruction fetch)
static __device__ __noinline__ double div_z(double x, double d) {
  if (x == 0.0 && d == d && d != 0.0) {
    const long long sx = __double_as_longlong(x) ^ __double_as_longlong(d);
    return __longlong_as_double(sx & (long long)0x8000000000000000ull);
  }
  return x / d;
}

// t^k for k = 0..6, correctly rounded (double-double running product, one final rounding): the
// time-polynomial systems below are badly conditioned, so a 1-2 ulp difference in a power (CUDA's
// pow vs the host libm's, which is correctly rounded for these arguments) would show up as 1e-7 in
// the fitted accelerations.
__device__ __forceinline__ double powi(double t, int k) {
  if (k == 0) return 1.0;
  double hi = t, lo = 0.0;
  for (int q = 1; q < k; q++) {
    const double p = hi * t;
    double e = fma(hi, t, -p);
    e = fma(lo, t, e);
    const double s = p + e;
    lo = e - (s - p);
    hi = s;
  }
  return hi + lo;
}

// pw[k] = powi(t, k) for k = 0..6 from ONE running product: powi's chain for t^k passes through the states of
// t^2 .. t^(k-1), so rounding every intermediate state gives bit-identical values at a sixth of the multiplications
__device__ __forceinline__ void powi_all(double t, double pw[7]) {
  pw[0] = 1.0;
  double hi = t, lo = 0.0;
  pw[1] = hi + lo;
#pragma unroll
  for (int q = 1; q < 6; q++) {
    const double p = hi * t;
    double e = fma(hi, t, -p);
    e = fma(lo, t, e);
    const double s = p + e;
    lo = e - (s - p);
    hi = s;
    pw[q + 1] = hi + lo;
  }
}

}  // namespace go1
