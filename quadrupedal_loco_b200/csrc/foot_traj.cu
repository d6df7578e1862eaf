// foot_traj.cu -- swing-foot trajectory of the step planner, one thread per planner.
//
// Replaces NLPClass::Foot_trajectory_solve_mod2 (NLP/src/NLP/NLPClass_sqp.cpp:2039-2358) and
// solve_AAA_inv2 (:3633-3645).  Compiled with -fmad=false and summing in the reference's order:
// the cubic fit is badly conditioned at some ticks (its first node comes within 1e-4 s of the
// mid-swing node), so the arithmetic is kept bit-identical to the CPU oracle.
#include <cuda_runtime.h>
#include "kernels.h"
#include "powi.cuh"

namespace go1 {

namespace {
constexpr int NS = 27;
constexpr int S_TS = 0, S_TX = 27, S_FX = 54, S_FY = 81, S_FZ = 108, S_BJX1 = 201;
}  // namespace

// ---------------------------------------------------------------------------------------------
// Swing-foot trajectory of the step planner: NLPClass::Foot_trajectory_solve_mod2
// (NLPClass_sqp.cpp:2039-2358) with solve_AAA_inv2 (:3633-3645); run after the step-timing tick
// of the same index.  Thread per instance, SoA.  The reference's whole-walk foot arrays shrink
// to a 32-double window (layout: include/go1mpc.h).  Stop-walking: FootKParams::lift0 / stop.
__device__ void gj_inverse4(double* a, double* r) {
  constexpr int n = 4;
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) r[i * n + j] = (i == j) ? 1.0 : 0.0;
  for (int k = 0; k < n; k++) {
    int piv = k;
    double best = fabs(a[k * n + k]);
    for (int i = k + 1; i < n; i++) if (fabs(a[i * n + k]) > best) { best = fabs(a[i * n + k]); piv = i; }
    if (piv != k)
      for (int j = 0; j < n; j++) {
        double t = a[k * n + j]; a[k * n + j] = a[piv * n + j]; a[piv * n + j] = t;
        t = r[k * n + j]; r[k * n + j] = r[piv * n + j]; r[piv * n + j] = t;
      }
    const double d = a[k * n + k];
    for (int j = 0; j < n; j++) { a[k * n + j] = a[k * n + j] / d; r[k * n + j] = r[k * n + j] / d; }
    for (int i = 0; i < n; i++) {
      if (i == k) continue;
      const double f = a[i * n + k];
      for (int j = 0; j < n; j++) { a[i * n + j] -= f * a[k * n + j]; r[i * n + j] -= f * r[k * n + j]; }
    }
  }
}

__global__ void __launch_bounds__(128) foot_traj_kernel(FootKParams P) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.B) return;
  const size_t B = (size_t)P.B;
  const double* S = P.state + b;
  double* F = P.foot + b;
#define ST(f) S[(size_t)(f) * B]
#define FS(f) F[(size_t)(f) * B]
  const double dt = P.dt, sw0 = P.stepwidth0;
  const int j = P.tick[b];
  if (j < 1) return;          // no tick for this planner
  const int bjx1 = (int)ST(S_BJX1);
  // _lift_height_ref(bjx1 - 1): FootStepInputs :71-75 (last two steps 0, the one before half) and the stop-walking
  // branch :2043-2048, which zeroes the steps ahead of the current one for good
  int lift0 = NS;
  if (P.lift0) {
    lift0 = (int)P.lift0[b];
    const bool stop = (P.stop && P.stop[b] != 0.0) || j > P.t_end;
    if (stop && bjx1 + 1 < lift0) { lift0 = bjx1 + 1; P.lift0[b] = (double)lift0; }
  }
  const int ks = bjx1 - 1;
  const double lift_h = (ks >= lift0 || ks >= NS - 2) ? 0.0 : (ks == NS - 3 ? P.lift_height / 2 : P.lift_height);
  const int bjxx = (int)P.out38[(size_t)27 * B + b];
  double pm1[6], pj[6], pm2[6], pm3[6], frz[6];
  for (int k = 0; k < 6; k++) { pm1[k] = FS(k); pj[k] = FS(6 + k); pm2[k] = FS(12 + k); pm3[k] = FS(18 + k); frz[k] = FS(24 + k); }
  double frz_s = FS(30), ry_lr = FS(31);
  double cur[6], nxt[6], vel[6] = {0, 0, 0, 0, 0, 0}, acc[6] = {0, 0, 0, 0, 0, 0};
  bool wrote_next[6] = {false, false, false, false, false, false};
  for (int k = 0; k < 6; k++) { cur[k] = pj[k]; nxt[k] = 0.0; }
  int right_support;
  // _footxyz_real: the step tables with (1,0) overwritten by -stepwidth(0) (:2050)
  auto fxr = [&](int r, int k) -> double {
    k = k < 0 ? 0 : (k > NS - 1 ? NS - 1 : k);
    return r == 0 ? ST(S_FX + k) : (r == 1 ? (k == 0 ? -sw0 : ST(S_FY + k)) : ST(S_FZ + k));
  };
  if (bjx1 >= 2 && bjx1 <= NS) {
    const double tx1 = ST(S_TX + bjx1 - 1), ts1 = ST(S_TS + bjx1 - 1);
    const int s = (int)round(tx1 / dt);
    if ((double)s != frz_s) {
      const int back = j - (s - 2);
      for (int k = 0; k < 6; k++) frz[k] = (back <= 1) ? pm1[k] : (back == 2 ? pm2[k] : pm3[k]);
      frz_s = (double)s;
    }
    const bool left_support = (bjx1 % 2 == 0);
    const int so = left_support ? 3 : 0, wo = left_support ? 0 : 3;
    right_support = left_support ? 0 : 1;
    for (int k = 0; k < 3; k++) { cur[so + k] = frz[so + k]; nxt[so + k] = frz[so + k]; wrote_next[so + k] = true; }
    if ((j + 1 - s) * dt < 0.2 * ts1) {
      right_support = 2;
      for (int k = 0; k < 3; k++) { cur[wo + k] = frz[wo + k]; nxt[wo + k] = frz[wo + k]; wrote_next[wo + k] = true; }
    } else {
      const double t_des = (j + 1 - s + 1) * dt;
      const double td1 = 0.2 * ts1;
      const double tp[3] = {t_des - dt, (td1 + ts1) / 2 + 0.0001, ts1};
      if (fabs(t_des - ts1) <= (+0.0005)) {
        for (int k = 0; k < 3; k++) { cur[wo + k] = fxr(k, bjxx); nxt[wo + k] = fxr(k, bjxx); wrote_next[wo + k] = true; }
      } else {
        double A[16], Ai[16];
        for (int r = 0; r < 3; r++) { A[4 * r] = powi(tp[r], 3); A[4 * r + 1] = powi(tp[r], 2); A[4 * r + 2] = powi(tp[r], 1); A[4 * r + 3] = 1; }
        A[12] = 3 * powi(tp[2], 2); A[13] = 2 * powi(tp[2], 1); A[14] = powi(tp[2], 0); A[15] = 0;
        gj_inverse4(A, Ai);
        const double tap[4] = {powi(t_des, 3), powi(t_des, 2), powi(t_des, 1), 1};
        const double tav[4] = {3 * powi(t_des, 2), 2 * powi(t_des, 1), 1, 0};
        const double taa[4] = {6 * powi(t_des, 1), 2, 0, 0};
        if ((j + 1 - s) * dt < td1 + dt) ry_lr = (fxr(1, bjxx) + fxr(1, bjxx - 2)) / 2;
        for (int k = 0; k < 3; k++) {
          double plan[4];
          plan[0] = pm1[wo + k];
          if (k == 0) plan[1] = (fxr(0, bjxx - 2) + fxr(0, bjxx)) / 2;
          else if (k == 1) plan[1] = ry_lr;
          else plan[1] = fmax(fxr(2, bjxx - 2), fxr(2, bjxx)) + lift_h;
          plan[2] = fxr(k, bjxx);
          plan[3] = 0;
          double co[4];
          for (int r = 0; r < 4; r++) { double a_ = 0.0; for (int q = 0; q < 4; q++) a_ += Ai[4 * r + q] * plan[q]; co[r] = a_; }
          double p_ = 0.0, v_ = 0.0, a2 = 0.0;
          for (int q = 0; q < 4; q++) { p_ += tap[q] * co[q]; v_ += tav[q] * co[q]; a2 += taa[q] * co[q]; }
          cur[wo + k] = p_; vel[wo + k] = v_; acc[wo + k] = a2;
          nxt[wo + k] = cur[wo + k] + dt * vel[wo + k];
          wrote_next[wo + k] = true;
        }
      }
    }
  } else {
    right_support = 2;
    cur[1] = -sw0;
    cur[4] = sw0;
  }
  double* O = P.out18 + b;
  for (int k = 0; k < 6; k++) { O[(size_t)k * B] = cur[k]; O[(size_t)(6 + k) * B] = vel[k]; O[(size_t)(12 + k) * B] = acc[k]; }
  if (P.right_support) P.right_support[b] = right_support;
  const double init[6] = {0, -sw0, 0, 0, sw0, 0};
  for (int k = 0; k < 6; k++) {
    FS(18 + k) = pm2[k]; FS(12 + k) = pm1[k]; FS(k) = cur[k];
    FS(6 + k) = wrote_next[k] ? nxt[k] : init[k];
    FS(24 + k) = frz[k];
  }
  FS(30) = frz_s; FS(31) = ry_lr;
#undef ST
#undef FS
}

cudaError_t foot_traj_launch(FootKParams P, cudaStream_t st) {
  foot_traj_kernel<<<(P.B + 127) / 128, 128, 0, st>>>(P);
  return cudaGetLastError();
}

}  // namespace go1
