// ref_interp.cu -- 40 Hz -> 100 Hz reference interpolation of rt_mpc_qp, batched.
//
// Replaces PRMPCClass::XGetSolution_position_mod3 (RT/src/FastMPC/PRMPCClass.cpp:1170-1261) as gait_fast.cpp:131-138
// calls it, once per interpolated quantity (CoM, CoM acceleration, ZMP, DCM): a cubic through four consecutive 40 Hz
// samples (t = -dt, 0, dt, 2 dt), evaluated at the horizon's nh instants -- position / velocity / acceleration at
// the first, positions at the nh - 1 later ones: the reference rows of the body-inclination MPC.  _AAA_inv_mod
// (solve_AAA_inv_mod1, :1344-1361) depends only on dt and is built once on the host (api.cu).
// One thread per (robot, quantity) item; SoA [f * B + b].  The reference's operation order, no FMA contraction
// (-fmad=false), monomials by the correctly rounded powi -- bit-identical to the oracle.
// NOT YET RUN ON HARDWARE at the end of round 1 (written after the GPU budget was spent): tests/test_zz_ref_interp.py.
#include "kernels.h"
#include "powi.cuh"

namespace go1 {

namespace {
__device__ __forceinline__ double row_inv_temp(const double row[4], const double* inv, const double temp[4]) {
  double v[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 4; k++) acc += row[k] * inv[4 * k + j];
    v[j] = acc;
  }
  double acc = 0.0;
#pragma unroll
  for (int k = 0; k < 4; k++) acc += v[k] * temp[k];
  return acc;
}
}  // namespace

__global__ void __launch_bounds__(256) ref_interp_kernel(RefInterpParams P) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.B) return;
  const size_t B = (size_t)P.B;
  const int rows = 9 + 3 * (P.nh - 1);
  const int walktime = P.walktime[b];
  if (!(walktime <= P.t_end_footstep)) {
    for (int r = 0; r < rows; r++) P.out[r * B + b] = 0.0;
    return;
  }
  double s[12];
#pragma unroll
  for (int k = 0; k < 12; k++) s[k] = P.samples[k * B + b];     // in1 xyz | in2 xyz | ref xyz | ref2 xyz
  for (int jx = 0; jx < P.nh; jx++) {
    const double t = walktime * P.dt_sample + jx * P.dt_sample;
    const double t2 = powi(t, 2), t3 = powi(t, 3);
    const double p[4] = {t3, t2, t, 1.0};
    const double v[4] = {3 * t2, 2 * t, 1.0, 0.0};
    const double a[4] = {6 * t, 2.0, 0.0, 0.0};
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const double temp[4] = {s[c], s[3 + c], s[6 + c], s[9 + c]};
      if (jx == 0) {
        P.out[(size_t)c * B + b] = row_inv_temp(p, P.inv, temp);
        P.out[(size_t)(3 + c) * B + b] = row_inv_temp(v, P.inv, temp);
        P.out[(size_t)(6 + c) * B + b] = row_inv_temp(a, P.inv, temp);
      } else {
        P.out[(size_t)(8 + 3 * jx - 2 + c) * B + b] = row_inv_temp(p, P.inv, temp);
      }
    }
  }
}

cudaError_t ref_interp_launch(RefInterpParams P, cudaStream_t st) {
  ref_interp_kernel<<<(P.B + 255) / 256, 256, 0, st>>>(P);
  return cudaGetLastError();
}

}  // namespace go1
