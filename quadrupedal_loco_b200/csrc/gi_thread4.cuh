// gi_thread4.cuh -- the step-timing QP (n = 4, p = 1, m = 24), one thread per problem, solver state in REGISTERS.
//
// Same algorithm, tie rules, tolerances and operation order as GiThread<4, 1, 24> (gi_thread.cuh), i.e.
// Eigen::QP::solve_quadprog2 (RT/src/utils/EiQuadProg/EiQuadProg.cpp:172-491), add_constraint :30-93,
// delete_constraint :95-170, helpers EiQuadProg.hpp:100-134 -- the including translation unit is compiled with
// -fmad=false, so for identical inputs the arithmetic is bit-identical to the CPU oracle.
//
// What changes is only how the state is addressed.  The generic solver indexes J, R, d, u, A ... with the
// run-time size of the working set, which sends all of them (and the 24 slacks, 26-entry dual / index vectors
// sized for m + p constraints) to per-thread local memory: 1.3 KB per thread, the planner kernel's largest source
// of long-scoreboard stalls and DRAM traffic.  Here
//   * the working set never holds more than n = 4 constraints (+ the candidate), so u, A and their copies are 5 long;
//   * every loop runs over its compile-time range and the run-time bound is a predicate (`j >= iq`), so every
//     array index is a compile-time constant and the arrays live in registers;
//   * the slacks are not stored: step 1 and step 2 (cpp:282-342) are one scan that keeps the running arg-min and the
//     candidate's column; the re-scan after a degenerate add (where the reference reads its stale slacks, cpp:461)
//     recomputes them from the restored x, which is the x they were computed from -- the same values.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include "powi.cuh"

namespace go1 {

// Inequality rows as dense arrays (CI 4 x 24 column-major, ci0): the generic way to hand them to GiThread4.
struct GiRowsArray {
  const double* CI;
  const double* ci0;
  __device__ __forceinline__ double slack(int i, const double* x) const {
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 4; k++) acc += CI[i * 4 + k] * x[k];
    return acc + ci0[i];
  }
  __device__ __forceinline__ void column(int i, double* np) const {
#pragma unroll
    for (int k = 0; k < 4; k++) np[k] = CI[i * 4 + k];
  }
  __device__ __forceinline__ double rhs(int i) const { return ci0[i]; }
};

struct GiThread4 {
  static constexpr int N = 4, P = 1, M = 24;
  static constexpr double EPS = 2.220446049250313e-16;
  double J[16], R[16];            // column-major: J[j*4 + k] = J(k, j)
  double z[4], r[5], d[4], np[4], u[5], x_old[4], u_old[5];
  int A[5], A_old[5];
  unsigned inA, excl;
  int it_outer, it_add, it_drop, it_degen, iq;
  double f_value;

  __device__ __noinline__ static double hyp(double a, double b) {
    double a1 = fabs(a), b1 = fabs(b), t;
    if (a1 > b1) { t = div_z(b1, a1); return a1 * sqrt(1.0 + t * t); }
    if (b1 > a1) { t = div_z(a1, b1); return b1 * sqrt(1.0 + t * t); }
    return a1 * sqrt(2.0);
  }
  __device__ __forceinline__ static double dot4(const double* a, const double* b) {
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++) acc += a[i] * b[i];
    return acc;
  }
  __device__ __forceinline__ void compute_d() {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < 4; k++) acc += J[j * 4 + k] * np[k];
      d[j] = acc;
    }
  }
  __device__ __forceinline__ void update_z() {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < 4; j++) if (j >= iq) acc += J[j * 4 + k] * d[j];
      z[k] = acc;
    }
  }
  __device__ __forceinline__ void update_r() {
#pragma unroll
    for (int i = 0; i < 4; i++) if (i < iq) r[i] = d[i];
#pragma unroll
    for (int i = 3; i >= 0; i--)
      if (i < iq) {
        r[i] = div_z(r[i], R[i * 4 + i]);
        const double ri = r[i];
#pragma unroll
        for (int t = 0; t < i; t++) r[t] -= ri * R[i * 4 + t];
      }
  }

  __device__ __forceinline__ bool add_constraint(double& R_norm) {
#pragma unroll
    for (int j = 3; j >= 1; j--)
      if (j >= iq + 1) {
        double cc = d[j - 1], ss = d[j];
        const double h = hyp(cc, ss);
        if (h != 0.0) {
          d[j] = 0.0;
          ss = div_z(ss, h); cc = div_z(cc, h);
          if (cc < 0.0) { cc = -cc; ss = -ss; d[j - 1] = -h; } else d[j - 1] = h;
          const double xny = div_z(ss, 1.0 + cc);
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const double t1 = J[(j - 1) * 4 + k], t2 = J[j * 4 + k];
            const double a = t1 * cc + t2 * ss;
            J[(j - 1) * 4 + k] = a;
            J[j * 4 + k] = xny * (t1 + a) - t2;
          }
        }
      }
    iq++;
    double dq = 0.0;
#pragma unroll
    for (int c = 0; c < 4; c++)
      if (c == iq - 1) {
#pragma unroll
        for (int i = 0; i < 4; i++) if (i < iq) R[c * 4 + i] = d[i];
        dq = d[c];
      }
    if (fabs(dq) <= EPS * R_norm) return false;
    R_norm = fmax(R_norm, fabs(dq));
    return true;
  }

  __device__ __forceinline__ bool delete_constraint(int l) {
    int qq = -1;
#pragma unroll
    for (int i = 3; i >= P; i--) if (i < iq && A[i] == l) qq = i;      // lowest matching slot, as the forward scan with break
    if (qq < 0) return false;
#pragma unroll
    for (int i = 0; i < 3; i++)
      if (i >= qq && i < iq - 1) {
        A[i] = A[i + 1]; u[i] = u[i + 1];
#pragma unroll
        for (int k = 0; k < 4; k++) R[i * 4 + k] = R[(i + 1) * 4 + k];
      }
    // A[iq-1] = A[iq]; u[iq-1] = u[iq]; A[iq] = 0; u[iq] = 0; R(:, iq-1)[0..iq) = 0
#pragma unroll
    for (int i = 1; i <= 4; i++)
      if (i == iq) {
        A[i - 1] = A[i]; u[i - 1] = u[i]; A[i] = 0; u[i] = 0.0;
#pragma unroll
        for (int j = 0; j < 4; j++) if (j < iq) R[(i - 1) * 4 + j] = 0.0;
      }
    iq--;
    if (iq == 0) return true;
#pragma unroll
    for (int j = 0; j < 3; j++)
      if (j >= qq && j < iq) {
        double cc = R[j * 4 + j], ss = R[j * 4 + j + 1];
        const double h = hyp(cc, ss);
        if (h != 0.0) {
          cc = div_z(cc, h); ss = div_z(ss, h);
          R[j * 4 + j + 1] = 0.0;
          if (cc < 0.0) { R[j * 4 + j] = -h; cc = -cc; ss = -ss; } else R[j * 4 + j] = h;
          const double xny = div_z(ss, 1.0 + cc);
#pragma unroll
          for (int k = j + 1; k < 4; k++)
            if (k < iq) {
              const double t1 = R[k * 4 + j], t2 = R[k * 4 + j + 1];
              const double a = t1 * cc + t2 * ss;
              R[k * 4 + j] = a;
              R[k * 4 + j + 1] = xny * (t1 + a) - t2;
            }
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const double t1 = J[j * 4 + k], t2 = J[(j + 1) * 4 + k];
            const double a = t1 * cc + t2 * ss;
            J[j * 4 + k] = a;
            J[(j + 1) * 4 + k] = xny * (a + t1) - t2;
          }
        }
      }
    return true;
  }

  // Steps 1 and 2 in one pass over the 24 constraints (cpp:282-342).  first = true: step 1 (psi, reset of ss / ip)
  // followed by step 2; first = false: the step-2 re-scan after a degenerate add (ss keeps its value, cpp:461).
  // Returns psi; ss / ip / s_ip / np are updated when a better candidate is found.
  template <class Rows>
  __device__ __forceinline__ double scan(const Rows& rows, const double* x, bool first,
                                         double& ss, int& ip, double& s_ip) {
    double psi = 0.0;
    if (first) { ss = 0.0; ip = 0; }
#pragma unroll
    for (int i = 0; i < M; i++) {
      const double sum = rows.slack(i, x);      // i is a compile-time constant after unrolling
      psi += (sum < 0.0) ? sum : 0.0;           // = fmin(0.0, sum) for every input that matters: NaN gives 0 either way, and the
                                                // sign of a zero cannot change psi; fmin's IEEE handling costs 5 instructions more per row
      if (sum < ss && !((inA >> i) & 1u) && !((excl >> i) & 1u)) {
        ss = sum; ip = i; s_ip = sum;
        rows.column(i, np);
      }
    }
    // nothing found: ss / ip keep their values, as in the reference's scan
    return psi;
  }

  // G, CE (4 x 1), CI (4 x 24) column-major.  x: in/out.  Returns the status code of go1mpc.h.
  __device__ int solve(const double* G, const double* g0, const double* CE, const double* ce0,
                       const double* CI, const double* ci0, double* x, int cap) {
    GiRowsArray rows{CI, ci0};
    return solve_rows(G, g0, CE, ce0, rows, x, cap);
  }
  // the same with the inequality rows behind a policy: slack(i, x) = CI(:, i)' x + ci0(i) summed in dot-product order,
  // column(i, np) = CI(:, i), rhs(i) = ci0(i); i is a compile-time constant wherever the solver calls them
  template <class Rows>
  __device__ int solve_rows(const double* G, const double* g0, const double* CE, const double* ce0,
                            const Rows& rows, double* x, int cap) {
    const double inf = CUDART_INF;
    double L[16], y[4];
    it_outer = it_add = it_drop = it_degen = 0; iq = 0; inA = 0u; excl = 0u;
#pragma unroll
    for (int i = 0; i < 5; i++) { A[i] = 0; A_old[i] = 0; u[i] = 0.0; u_old[i] = 0.0; r[i] = 0.0; }
    double c1 = 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++) c1 += G[i * 4 + i];
#pragma unroll
    for (int i = 0; i < 16; i++) L[i] = G[i];
    // unblocked left-looking lower Cholesky (Eigen LLT, n < 32 path)
#pragma unroll
    for (int k = 0; k < 4; k++) {
      double xx = L[k * 4 + k], sq = 0.0;
#pragma unroll
      for (int j = 0; j < k; j++) { const double v = L[j * 4 + k]; sq += v * v; }
      if (k > 0) xx -= sq;
      if (xx <= 0.0) { f_value = inf; return 1; }
      xx = sqrt(xx);
      L[k * 4 + k] = xx;
#pragma unroll
      for (int i = k + 1; i < 4; i++) {
        double t = 0.0;
#pragma unroll
        for (int j = 0; j < k; j++) t += L[j * 4 + i] * L[j * 4 + k];
        double v = L[k * 4 + i];
        if (k > 0) v -= t;
        L[k * 4 + i] = div_z(v, xx);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) d[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 16; i++) R[i] = 0.0;
    double R_norm = 1.0;
    // J = L^-T
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
      for (int i = 3; i >= 0; i--) {
        double t = 0.0;
#pragma unroll
        for (int k = i + 1; k < 4; k++) t += L[i * 4 + k] * J[c * 4 + k];
        double rhs = (i == c) ? 1.0 : 0.0;
        if (i < 3) rhs -= t;
        J[c * 4 + i] = div_z(rhs, L[i * 4 + i]);
      }
    double c2 = 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++) c2 += J[i * 4 + i];
    // x = -G^-1 g0
#pragma unroll
    for (int i = 0; i < 4; i++) y[i] = g0[i];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      y[i] = div_z(y[i], L[i * 4 + i]);
      const double yi = y[i];
#pragma unroll
      for (int k = i + 1; k < 4; k++) y[k] -= yi * L[i * 4 + k];
    }
#pragma unroll
    for (int i = 3; i >= 0; i--) {
      double t = 0.0;
#pragma unroll
      for (int k = i + 1; k < 4; k++) t += L[i * 4 + k] * y[k];
      double rhs = y[i];
      if (i < 3) rhs -= t;
      y[i] = div_z(rhs, L[i * 4 + i]);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) x[i] = -y[i];
    f_value = 0.5 * dot4(g0, x);
    int status = 0;
    // the equality constraint (cpp:236-276)
    {
      bool allzero = true;
#pragma unroll
      for (int k = 0; k < 4; k++) if (!(fabs(CE[k]) <= 1e-12)) allzero = false;
      if (!allzero) {
#pragma unroll
        for (int k = 0; k < 4; k++) np[k] = CE[k];
        compute_d(); update_z(); update_r();
        double t2 = 0.0;
        if (fabs(dot4(z, z)) > EPS) t2 = (-dot4(np, x) - ce0[0]) / dot4(z, np);
#pragma unroll
        for (int k = 0; k < 4; k++) x[k] += t2 * z[k];
        u[0] = t2;                      // iq = 0 here
        f_value += 0.5 * (t2 * t2) * dot4(z, np);
        A[0] = -1;
        if (!add_constraint(R_norm)) return 5;
      }
    }
    enum { PH_L1, PH_L2, PH_L2A };
    int phase = PH_L1, ip = 0, l = 0, passes = 0;
    double ss = 0.0, s_ip = 0.0;
    for (;;) {
      if (phase == PH_L1) {
        it_outer++;
#pragma unroll
        for (int i = P; i < 4; i++) if (i < iq) inA |= 1u << A[i];
        excl = 0u;
        const double psi = scan(rows, x, true, ss, ip, s_ip);
        if (fabs(psi) <= M * EPS * c1 * c2 * 100.0) break;
#pragma unroll
        for (int i = 0; i < 4; i++) if (i < iq) { u_old[i] = u[i]; A_old[i] = A[i]; }
#pragma unroll
        for (int k = 0; k < 4; k++) x_old[k] = x[k];
        phase = PH_L2;
      } else if (phase == PH_L2) {
        (void)scan(rows, x, false, ss, ip, s_ip);
      }
      if (phase == PH_L2) {
        if (ss >= 0.0) break;
#pragma unroll
        for (int i = 0; i < 5; i++) if (i == iq) { u[i] = 0.0; A[i] = ip; }
        phase = PH_L2A;
      }
      if (++passes > cap) { status = 3; break; }
      compute_d(); update_z(); update_r();
      l = 0;
      double t1 = inf, t2;
#pragma unroll
      for (int k = P; k < 4; k++)
        if (k < iq && r[k] > 0.0) { const double tmp = div_z(u[k], r[k]); if (tmp < t1) { t1 = tmp; l = A[k]; } }
      if (fabs(dot4(z, z)) > EPS) t2 = -s_ip / dot4(z, np); else t2 = inf;
      const double t = fmin(t1, t2);
      if (t >= inf) { status = 2; f_value = inf; break; }
      double uiq = 0.0;
#pragma unroll
      for (int i = 0; i < 5; i++) if (i == iq) uiq = u[i];
      if (t2 >= inf) {
#pragma unroll
        for (int k = 0; k < 4; k++) if (k < iq) u[k] -= t * r[k];
#pragma unroll
        for (int i = 0; i < 5; i++) if (i == iq) u[i] += t;
        inA &= ~(1u << l);
        if (!delete_constraint(l)) { status = 3; break; }
        it_drop++;
        continue;
      }
      {
        const double zn = dot4(z, np);
#pragma unroll
        for (int k = 0; k < 4; k++) x[k] += t * z[k];
        f_value += t * zn * (0.5 * t + uiq);
      }
#pragma unroll
      for (int k = 0; k < 4; k++) if (k < iq) u[k] -= t * r[k];
#pragma unroll
      for (int i = 0; i < 5; i++) if (i == iq) u[i] += t;
      if (t == t2) {
        if (!add_constraint(R_norm)) {
          it_degen++;
          excl |= 1u << ip;
          if (!delete_constraint(ip)) { status = 3; break; }
          inA = 0u;
#pragma unroll
          for (int i = 0; i < 4; i++) if (i < iq) { A[i] = A_old[i]; if (A[i] >= 0) inA |= 1u << A[i]; u[i] = u_old[i]; }
#pragma unroll
          for (int k = 0; k < 4; k++) x[k] = x_old[k];
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
      {
        // s[ip] at the new x: ip is run-time, the column is still in np
        s_ip = dot4(np, x);
        double cv = 0.0;
#pragma unroll
        for (int i = 0; i < M; i++) if (i == ip) cv = rows.rhs(i);
        s_ip = s_ip + cv;
      }
    }
    if (status == 0) {
#pragma unroll
      for (int i = 0; i < 4; i++) if (x[i] != x[i]) status = 4;
    }
    return status;
  }
};

}  // namespace go1
