// Body-inclination MPC tick with the slow-changing part of the record RESIDENT on the device.
//
// What a controller really changes per tick is tick / theta / measured theta and the 9 reference rows
// (PRMPCClass::body_theta_mpc's arguments, RT/src/FastMPC/PRMPCClass.cpp:379-395); the step table _tx (:174-178)
// only moves when the planner re-times a step and _V_ini (:258-261) is the previous tick's own result.  The
// pipelined host entry therefore keeps tx and the previous output record in HBM and moves 9+9nh doubles up and
// 20 doubles down per instance instead of 36+11nh up and 19+2nh down.
//
// Two copy kernels around the unchanged body tick (thread per double, coalesced on the written side):
//   body_record_expand_kernel : full input records from tx_d, the tick records and the resident x (warm start)
//   body_record_pack_kernel   : out14 | theta | cost of the resident output records -> 20-double tick results
#include "kernels.h"

namespace go1 {

__global__ void __launch_bounds__(256) body_record_expand_kernel(int B, int nh, int in_stride, int tick_stride, int out_stride,
                                                                 const double* __restrict__ tx, const double* __restrict__ tick_in,
                                                                 const double* __restrict__ out_res, double* __restrict__ rec) {
  const size_t total = (size_t)B * in_stride;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t b = idx / in_stride;
    const int j = (int)(idx - b * in_stride);
    double v = 0.0;
    if (j < 27) v = tx[b * 28 + j];
    else if (j < 36) v = tick_in[b * tick_stride + (j - 27)];
    else if (j < 36 + 2 * nh) v = out_res[b * out_stride + 18 + (j - 36)];
    else if (j < 36 + 11 * nh) v = tick_in[b * tick_stride + 9 + (j - 36 - 2 * nh)];
    rec[idx] = v;
  }
}

__global__ void __launch_bounds__(256) body_record_pack_kernel(int B, int nh, int out_stride, const double* __restrict__ out_res,
                                                               double* __restrict__ tick_out) {
  const size_t total = (size_t)B * 20;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t b = idx / 20;
    const int j = (int)(idx - b * 20);
    double v = 0.0;
    if (j < 18) v = out_res[b * out_stride + j];
    else if (j == 18) v = out_res[b * out_stride + 18 + 2 * nh];
    tick_out[idx] = v;
  }
}

// Compact per-robot result of one control tick (planner + body MPC): what a controller consumes per tick plus the
// solver statuses, 12 doubles per robot, instance-major (one 96-byte row per robot: the unit of the once-per-batch
// gather to rank 0 and of the device -> host read of the e2e path).
//   0..2  CoM position x, y, z (out38 rows 0..2)      3..4  body roll / pitch (out14[0], out14[1])
//   5..6  body torques (out14[2], out14[3])            7..8  next footstep x, y (out38 rows 29, 31)
//   9     step period ts (out38 row 35)                10    planner: status of its last SQP solve (-1: none ran)
//   11    body MPC: QP status
__global__ void __launch_bounds__(256) compact_pack_kernel(int B, int nh, int out_stride, int diag_stride, const double* __restrict__ out38,
                                                           const int* __restrict__ step_diag, const double* __restrict__ body_out,
                                                           const int* __restrict__ body_diag, double* __restrict__ compact) {
  const size_t Bs = (size_t)B;
  for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < Bs; b += (size_t)gridDim.x * blockDim.x) {
    double r[12];
    r[0] = out38[0 * Bs + b]; r[1] = out38[1 * Bs + b]; r[2] = out38[2 * Bs + b];
    const double2* bo = reinterpret_cast<const double2*>(body_out + b * out_stride);
    const double2 o01 = bo[0], o23 = bo[1];
    r[3] = o01.x; r[4] = o01.y; r[5] = o23.x; r[6] = o23.y;
    r[7] = out38[29 * Bs + b]; r[8] = out38[31 * Bs + b]; r[9] = out38[35 * Bs + b];
    int sst = -1;
    if (step_diag) {
      const int ns = step_diag[4 * Bs + b];
      if (ns > 0) sst = step_diag[(size_t)(STEP_DIAG_HEAD + (ns - 1 < STEP_MAX_SQP ? ns - 1 : STEP_MAX_SQP - 1) * STEP_DIAG_PER) * Bs + b];
    }
    r[10] = (double)sst;
    r[11] = body_diag ? (double)body_diag[b * diag_stride] : 0.0;
    double2* c2 = reinterpret_cast<double2*>(compact + b * 12);
#pragma unroll
    for (int k = 0; k < 6; k++) c2[k] = make_double2(r[2 * k], r[2 * k + 1]);
  }
}

static int copy_grid(size_t total, int sms) {
  size_t g = (total + 255) / 256;
  const size_t cap = (size_t)sms * 8;
  return (int)(g < cap ? (g ? g : 1) : cap);
}

cudaError_t body_record_expand_launch(int B, int nh, int in_stride, int tick_stride, int out_stride, const double* tx,
                                      const double* tick_in, const double* out_res, double* rec, int sms, cudaStream_t st) {
  body_record_expand_kernel<<<copy_grid((size_t)B * in_stride, sms), 256, 0, st>>>(B, nh, in_stride, tick_stride, out_stride, tx, tick_in,
                                                                                out_res, rec);
  return cudaGetLastError();
}
cudaError_t body_record_pack_launch(int B, int nh, int out_stride, const double* out_res, double* tick_out, int sms, cudaStream_t st) {
  body_record_pack_kernel<<<copy_grid((size_t)B * 20, sms), 256, 0, st>>>(B, nh, out_stride, out_res, tick_out);
  return cudaGetLastError();
}

cudaError_t compact_pack_launch(int B, int nh, int out_stride, int diag_stride, const double* out38, const int* step_diag,
                                const double* body_out, const int* body_diag, double* compact, int sms, cudaStream_t st) {
  compact_pack_kernel<<<copy_grid((size_t)B, sms), 256, 0, st>>>(B, nh, out_stride, diag_stride, out38, step_diag, body_out, body_diag, compact);
  return cudaGetLastError();
}

}  // namespace go1
