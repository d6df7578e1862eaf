// powi.cuh -- correctly rounded small integer powers (shared by the time-polynomial fits).
#pragma once
#include <cuda_runtime.h>

namespace go1 {

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
