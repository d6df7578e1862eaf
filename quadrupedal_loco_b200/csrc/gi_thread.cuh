// gi_thread.cuh -- Goldfarb-Idnani dual active-set QP, one THREAD per problem, compile-time sizes.
//
// For the tiny QPs of the step-timing SQP (n = 4, p = 1, m = 24) a warp per problem would idle
// 28 of 32 lanes, so here every thread carries a whole problem in registers / local memory
// (local memory is interleaved per thread, so a warp's accesses coalesce) and 32 independent
// problems advance per warp instruction; the price is divergence when their iteration counts
// differ.  Algorithm, tie rules and tolerances: Eigen::QP::solve_quadprog2
// (RT/src/utils/EiQuadProg/EiQuadProg.cpp:172-491), add_constraint :30-93, delete_constraint
// :95-170, helpers EiQuadProg.hpp:100-134.  The translation unit that includes this header is
// compiled with -fmad=false and sums in the reference's order, so for identical inputs the QP
// arithmetic is bit-identical to the CPU oracle (only libm calls in the front-end differ).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

namespace go1 {

template <int N, int P, int M>
struct GiThread {
  static constexpr double EPS = 2.220446049250313e-16;
  double J[N * N], R[N * N];            // column-major
  double s[M], z[N], r[N + 1], d[N], np[N], u[M + P + 1], x_old[N], u_old[M + P + 1];
  int A[M + P + 1], A_old[M + P + 1];
  unsigned inA, excl;                   // bit i <-> inequality i (M <= 32)
  int it_outer, it_add, it_drop, it_degen, iq;
  double f_value;

  __device__ static double hyp(double a, double b) {
    double a1 = fabs(a), b1 = fabs(b), t;
    if (a1 > b1) { t = b1 / a1; return a1 * sqrt(1.0 + t * t); }
    if (b1 > a1) { t = a1 / b1; return b1 * sqrt(1.0 + t * t); }
    return a1 * sqrt(2.0);
  }
  __device__ void compute_d() {
    for (int j = 0; j < N; j++) { double acc = 0.0; for (int k = 0; k < N; k++) acc += J[j * N + k] * np[k]; d[j] = acc; }
  }
  __device__ void update_z() {
    for (int k = 0; k < N; k++) { double acc = 0.0; for (int j = iq; j < N; j++) acc += J[j * N + k] * d[j]; z[k] = acc; }
  }
  __device__ void update_r() {
    for (int i = 0; i < iq; i++) r[i] = d[i];
    for (int i = iq - 1; i >= 0; i--) {
      r[i] = r[i] / R[i * N + i];
      const double ri = r[i];
      for (int t = 0; t < i; t++) r[t] -= ri * R[i * N + t];
    }
  }
  __device__ static double dot(const double* a, const double* b) { double acc = 0.0; for (int i = 0; i < N; i++) acc += a[i] * b[i]; return acc; }

  __device__ bool add_constraint(double& R_norm) {
    for (int j = N - 1; j >= iq + 1; j--) {
      double cc = d[j - 1], ss = d[j];
      const double h = hyp(cc, ss);
      if (h == 0.0) continue;
      d[j] = 0.0;
      ss = ss / h; cc = cc / h;
      if (cc < 0.0) { cc = -cc; ss = -ss; d[j - 1] = -h; } else d[j - 1] = h;
      const double xny = ss / (1.0 + cc);
      for (int k = 0; k < N; k++) {
        const double t1 = J[(j - 1) * N + k], t2 = J[j * N + k];
        const double a = t1 * cc + t2 * ss;
        J[(j - 1) * N + k] = a;
        J[j * N + k] = xny * (t1 + a) - t2;
      }
    }
    iq++;
    for (int i = 0; i < iq; i++) R[(iq - 1) * N + i] = d[i];
    if (fabs(d[iq - 1]) <= EPS * R_norm) return false;
    R_norm = fmax(R_norm, fabs(d[iq - 1]));
    return true;
  }
  __device__ bool delete_constraint(int l) {
    int qq = -1;
    for (int i = P; i < iq; i++) if (A[i] == l) { qq = i; break; }
    if (qq < 0) return false;
    for (int i = qq; i < iq - 1; i++) {
      A[i] = A[i + 1]; u[i] = u[i + 1];
      for (int k = 0; k < N; k++) R[i * N + k] = R[(i + 1) * N + k];
    }
    A[iq - 1] = A[iq]; u[iq - 1] = u[iq]; A[iq] = 0; u[iq] = 0.0;
    for (int j = 0; j < iq; j++) R[(iq - 1) * N + j] = 0.0;
    iq--;
    if (iq == 0) return true;
    for (int j = qq; j < iq; j++) {
      double cc = R[j * N + j], ss = R[j * N + j + 1];
      const double h = hyp(cc, ss);
      if (h == 0.0) continue;
      cc = cc / h; ss = ss / h;
      R[j * N + j + 1] = 0.0;
      if (cc < 0.0) { R[j * N + j] = -h; cc = -cc; ss = -ss; } else R[j * N + j] = h;
      const double xny = ss / (1.0 + cc);
      for (int k = j + 1; k < iq; k++) {
        const double t1 = R[k * N + j], t2 = R[k * N + j + 1];
        const double a = t1 * cc + t2 * ss;
        R[k * N + j] = a;
        R[k * N + j + 1] = xny * (t1 + a) - t2;
      }
      for (int k = 0; k < N; k++) {
        const double t1 = J[j * N + k], t2 = J[(j + 1) * N + k];
        const double a = t1 * cc + t2 * ss;
        J[j * N + k] = a;
        J[(j + 1) * N + k] = xny * (a + t1) - t2;
      }
    }
    return true;
  }

  // G, CE (N x P), CI (N x M) column-major.  x: in/out.  Returns the status code of go1mpc.h.
  __device__ int solve(const double* G, const double* g0, const double* CE, const double* ce0,
                       const double* CI, const double* ci0, double* x, int cap) {
    const double inf = CUDART_INF;
    double L[N * N], y[N];
    it_outer = it_add = it_drop = it_degen = 0; iq = 0; inA = 0u; excl = 0u;
    for (int i = 0; i < M + P + 1; i++) { A[i] = 0; A_old[i] = 0; u[i] = 0.0; u_old[i] = 0.0; }
    for (int i = 0; i < N + 1; i++) r[i] = 0.0;
    double c1 = 0.0;
    for (int i = 0; i < N; i++) c1 += G[i * N + i];
    for (int i = 0; i < N * N; i++) L[i] = G[i];
    // unblocked left-looking lower Cholesky (Eigen LLT, n < 32 path)
    for (int k = 0; k < N; k++) {
      double xx = L[k * N + k], sq = 0.0;
      for (int j = 0; j < k; j++) { const double v = L[j * N + k]; sq += v * v; }
      if (k > 0) xx -= sq;
      if (xx <= 0.0) { f_value = inf; return 1; }
      xx = sqrt(xx);
      L[k * N + k] = xx;
      for (int i = k + 1; i < N; i++) {
        double t = 0.0;
        for (int j = 0; j < k; j++) t += L[j * N + i] * L[j * N + k];
        double v = L[k * N + i];
        if (k > 0) v -= t;
        L[k * N + i] = v / xx;
      }
    }
    for (int i = 0; i < N; i++) d[i] = 0.0;
    for (int i = 0; i < N * N; i++) R[i] = 0.0;
    double R_norm = 1.0;
    // J = L^-T
    for (int c = 0; c < N; c++)
      for (int i = N - 1; i >= 0; i--) {
        double t = 0.0;
        for (int k = i + 1; k < N; k++) t += L[i * N + k] * J[c * N + k];
        double rhs = (i == c) ? 1.0 : 0.0;
        if (i < N - 1) rhs -= t;
        J[c * N + i] = rhs / L[i * N + i];
      }
    double c2 = 0.0;
    for (int i = 0; i < N; i++) c2 += J[i * N + i];
    // x = -G^-1 g0
    for (int i = 0; i < N; i++) y[i] = g0[i];
    for (int i = 0; i < N; i++) {
      y[i] = y[i] / L[i * N + i];
      const double yi = y[i];
      for (int k = i + 1; k < N; k++) y[k] -= yi * L[i * N + k];
    }
    for (int i = N - 1; i >= 0; i--) {
      double t = 0.0;
      for (int k = i + 1; k < N; k++) t += L[i * N + k] * y[k];
      double rhs = y[i];
      if (i < N - 1) rhs -= t;
      y[i] = rhs / L[i * N + i];
    }
    for (int i = 0; i < N; i++) x[i] = -y[i];
    f_value = 0.5 * dot(g0, x);
    int status = 0;
    // equality constraints (cpp:236-276)
    for (int i = 0; i < P; i++) {
      const double* col = CE + i * N;
      bool allzero = true;
      for (int k = 0; k < N; k++) if (!(fabs(col[k]) <= 1e-12)) { allzero = false; break; }
      if (allzero) continue;
      for (int k = 0; k < N; k++) np[k] = col[k];
      compute_d(); update_z(); update_r();
      double t2 = 0.0;
      if (fabs(dot(z, z)) > EPS) t2 = (-dot(np, x) - ce0[i]) / dot(z, np);
      for (int k = 0; k < N; k++) x[k] += t2 * z[k];
      u[iq] = t2;
      for (int k = 0; k < iq; k++) u[k] -= t2 * r[k];
      f_value += 0.5 * (t2 * t2) * dot(z, np);
      A[i] = -i - 1;
      if (!add_constraint(R_norm)) return 5;
    }
    enum { PH_L1, PH_L2, PH_L2A };
    int phase = PH_L1, ip = 0, l = 0, passes = 0;
    double ss = 0.0;
    for (;;) {
      if (phase == PH_L1) {
        it_outer++;
        for (int i = P; i < iq; i++) inA |= 1u << A[i];
        ss = 0.0; ip = 0; excl = 0u;
        double psi = 0.0;
        for (int i = 0; i < M; i++) {
          const double sum = dot(CI + i * N, x) + ci0[i];
          s[i] = sum;
          psi += fmin(0.0, sum);
        }
        if (fabs(psi) <= M * EPS * c1 * c2 * 100.0) break;
        for (int i = 0; i < iq; i++) { u_old[i] = u[i]; A_old[i] = A[i]; }
        for (int k = 0; k < N; k++) x_old[k] = x[k];
        phase = PH_L2;
      }
      if (phase == PH_L2) {
        for (int i = 0; i < M; i++)
          if (s[i] < ss && !((inA >> i) & 1u) && !((excl >> i) & 1u)) { ss = s[i]; ip = i; }
        if (ss >= 0.0) break;
        for (int k = 0; k < N; k++) np[k] = CI[ip * N + k];
        u[iq] = 0.0; A[iq] = ip;
        phase = PH_L2A;
      }
      if (++passes > cap) { status = 3; break; }
      compute_d(); update_z(); update_r();
      l = 0;
      double t1 = inf, t2;
      for (int k = P; k < iq; k++) {
        if (r[k] > 0.0) { const double tmp = u[k] / r[k]; if (tmp < t1) { t1 = tmp; l = A[k]; } }
      }
      if (fabs(dot(z, z)) > EPS) t2 = -s[ip] / dot(z, np); else t2 = inf;
      const double t = fmin(t1, t2);
      if (t >= inf) { status = 2; f_value = inf; break; }
      if (t2 >= inf) {
        for (int k = 0; k < iq; k++) u[k] -= t * r[k];
        u[iq] += t;
        inA &= ~(1u << l);
        if (!delete_constraint(l)) { status = 3; break; }
        it_drop++;
        continue;
      }
      {
        const double zn = dot(z, np);
        for (int k = 0; k < N; k++) x[k] += t * z[k];
        f_value += t * zn * (0.5 * t + u[iq]);
      }
      for (int k = 0; k < iq; k++) u[k] -= t * r[k];
      u[iq] += t;
      if (t == t2) {
        if (!add_constraint(R_norm)) {
          it_degen++;
          excl |= 1u << ip;
          if (!delete_constraint(ip)) { status = 3; break; }
          inA = 0u;
          for (int i = 0; i < iq; i++) { A[i] = A_old[i]; if (A[i] >= 0) inA |= 1u << A[i]; u[i] = u_old[i]; }
          for (int k = 0; k < N; k++) x[k] = x_old[k];
          phase = PH_L2;
          continue;
        }
        it_add++;
        inA |= 1u << ip;
        phase = PH_L1;
        continue;
      }
      inA &= ~(1u << l);
      if (!delete_constraint(l)) { status = 3; break; }
      it_drop++;
      s[ip] = dot(CI + ip * N, x) + ci0[ip];
    }
    if (status == 0) for (int i = 0; i < N; i++) if (x[i] != x[i]) { status = 4; break; }
    return status;
  }
};

}  // namespace go1
