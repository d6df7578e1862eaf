// step_sqp.cu -- step-location / step-timing SQP tick, throughput mapping (one thread per planner), two launches.
//
// Replaces, for a batch of independent planners, one 40 Hz tick of
//   NLPClass::step_timing_opti_loop      NLP/src/NLP/NLPClass_sqp.cpp:693-1102
//   step_timing_object_function          :1144-1173
//   step_timing_constraints              :1175-1458
//   solve_stepping_timing / Solve        :1613-1653   (QP: n = 4, p = 1, m = 24; Eigen::QP::solve_quadprog2,
//                                        RT/src/utils/EiQuadProg/EiQuadProg.cpp:172-491, add/delete_constraint :30-170)
//   Indexfind                            :1105-1141
//   CoM_height_solve                     :2361-2473
// (NLP = unitree_ros/mosek_nlp_kmp).  Same contract as step_timing.cu (which keeps the warp-per-planner latency
// mapping for small batches): -fmad=false, the reference's operation order, IEEE divisions -- for identical inputs
// the QP arithmetic is bit-identical to the CPU oracle.
//
// What round 1's single kernel was bound by (profiles/r01_step_timing.md): 26.6 k instructions per warp, 12 k SASS
// (instruction fetch), 246 registers (6 warps / SM), a quarter of the instructions inside FP64-division subroutines.
// This file restructures the tick without changing a single rounding:
//   * THREE launches, each with the register budget / occupancy its phase wants and a fraction of the code:
//       step_sqp_kernel     front-end + the K SQP iterations (objective, constraint rows, the 4-variable QP)
//       step_height_kernel  the new step period and CoM_height_solve (7x7 inverse in registers: compute, few loads)
//       step_finish_kernel  write-back of step length / width / period and the step table, LIPM roll-out, feedback
//                           blend, integer step indices, outputs (memory-bound, few registers, many warps per SM)
//     the hand-over (the SQP point v, period index, k_yu, the height samples) travels through rows of the output buffer;
//   * the QP's Hessian G = 2 sym(SQ0) = diag(bbx, bby) (+) 2x2 does not depend on the SQP point: its Cholesky factor,
//     J = L^-T, trace terms and the stop tolerance are built ONCE per tick, exploiting the structural zeros (the
//     skipped terms are exact zeros, so every kept value is the dense algorithm's); the unconstrained minimiser
//     uses the same structure;
//   * divisions: a / d with a correctly rounded reciprocal y = RN(1 / d) at hand is q0 = a y, two FMA residual
//     corrections (Markstein: q2 = RN(q1 + y (a - q1 d)) is the correctly rounded quotient when q1 is faithful), i.e.
//     bit-identical to the IEEE division at a quarter of its instructions, and without the slow path CUDA's division
//     takes for zero numerators.  Reciprocals are shared: per pivot row of the 7x7 elimination, per diagonal entry of
//     R (kept beside R), per Cholesky pivot;
//   * Givens rotations that meet an exact zero (the three of the equality step meet two) take the closed form the
//     generic arithmetic reduces to (signed copy / sign flip), no hypot, no division;
//   * the 24 inequality rows live in shared memory ([slot][thread], conflict-free): rows 0-11 as right-hand sides
//     (their single coefficient is +-1), rows 12-23 as six +- pairs sharing two products; the candidate's column is
//     fetched by index after the scan instead of being carried through it.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <mutex>
#include "kernels.h"
#include "powi.cuh"

namespace go1 {

namespace {
constexpr int NS = 27;
constexpr int S_TS = 0, S_TX = 27, S_FX = 54, S_FY = 81, S_FZ = 108, S_LXX = 135, S_LYY = 162, S_FEED = 189, S_VARI = 195, S_END = 199, S_BJX1 = 201;
constexpr int I_EST = 0, I_RF = 6, I_LF = 8, I_CZ = 10, I_CAZ = 13, I_ZSC = 16, I_CVZ = 19;
constexpr double EPS = 2.220446049250313e-16;
constexpr int SQP_THREADS = 128;
// shared-memory slots per thread (doubles), [slot][thread]
constexpr int SM_BB = 0;        // 24 right-hand sides of the current SQP iteration
constexpr int SM_C2 = 24;       // 6 pairs: coefficient on tr1 (A-row of the even row)
constexpr int SM_M3A = 30;      // 6: CI(3, even row)
constexpr int SM_M3B = 36;      // 6: CI(3, odd row)
constexpr int SM_XOLD = 42;     // 4
constexpr int SM_UOLD = 46;     // 4
constexpr int SM_AOLD = 50;     // 1 (the packed working set, as a double)
constexpr int SM_U = 51;        // 4: duals of the working set
constexpr int SM_RINV = 55;     // 4: RN(1 / R(i, i))
constexpr int SM_R = 59;        // 10: R(i, j), i <= j, at j (j + 1) / 2 + i
constexpr int SM_SLOTS = 69;    // 70.6 KB per 128-thread block: three blocks per SM

// a / d, correctly rounded, given y = RN(1 / d) (d finite, non-zero): bit-identical to the IEEE division
__device__ __forceinline__ double div_rcp(double a, double d, double y) {
  const double q0 = a * y;
  const double e0 = fma(-q0, d, a);
  const double q1 = fma(e0, y, q0);
  const double e1 = fma(-q1, d, a);
  const double q2 = fma(e1, y, q1);
  return (a == 0.0) ? q0 : q2;
}
// (out of line, like hyp below: 11 call sites; inlined they are 1.4 k of the kernel's 7.3 k SASS instructions, and the kernel
// stalls on instruction fetch -- 202 -> 197 us per 65536 planners)
__device__ __noinline__ double div_full(double a, double d) { return div_rcp(a, d, __drcp_rn(d)); }

// EiQuadProg.hpp:100-118
__device__ __noinline__ double hyp(double a, double b) {
  const double a1 = fabs(a), b1 = fabs(b);
  if (a1 > b1) { const double t = div_full(b1, a1); return a1 * sqrt(1.0 + t * t); }
  if (b1 > a1) { const double t = div_full(a1, b1); return b1 * sqrt(1.0 + t * t); }
  return a1 * sqrt(2.0);
}

// The 4-variable QP of one SQP iteration; solver state in registers (every array index is a compile-time constant
// after unrolling, run-time bounds are predicates), rows / restore copies in shared memory.
struct Qp4 {
  double J[16];                    // column-major: J[j*4 + k] = J(k, j)
  double z[4], r[4], d[4], np[4], x[4];
  double un;                       // dual of the pending constraint (the reference's u[iq]); its index is the caller's ip
  unsigned Ap;                     // working set, 8 bits per slot: constraint index + 1 (equality slot -1 -> 0)
  // R (upper triangle), the reciprocals of its diagonal and the duals u live in shared memory (run-time indexed)
  unsigned inA, excl;
  int iq, it_outer, it_add, it_drop, it_degen;
  double f_value, R_norm;
  double* sm;                      // this thread's column of the shared-memory block (stride SQP_THREADS)

  __device__ __forceinline__ double& S(int slot) const { return sm[slot * SQP_THREADS]; }
  __device__ __forceinline__ double& Rs(int i, int j) const { return sm[(SM_R + j * (j + 1) / 2 + i) * SQP_THREADS]; }
  __device__ __forceinline__ int getA(int i) const { return (int)((Ap >> (8 * i)) & 0xffu) - 1; }
  __device__ __forceinline__ void setA(int i, int a) { Ap = (Ap & ~(0xffu << (8 * i))) | ((unsigned)(a + 1) << (8 * i)); }
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
        r[i] = div_rcp(r[i], Rs(i, i), S(SM_RINV + i));
        const double ri = r[i];
#pragma unroll
        for (int t = 0; t < i; t++) r[t] -= ri * Rs(t, i);
      }
  }

  // cpp:30-93.  A rotation whose (cc, ss) pair holds an exact zero reduces, in the reference's own arithmetic, to a
  // sign flip of column j (ss == 0: h = |cc|, cc/h = +-1, xny = +-0) or a signed copy (cc == 0: h = |ss|, ss/h = +-1,
  // xny = +-1); those closed forms are taken directly.
  // `pend`: index of the constraint being added (the reference's A[iq]); its dual is `un`
  __device__ __forceinline__ bool add_constraint(int pend) {
#pragma unroll
    for (int j = 3; j >= 1; j--)
      if (j >= iq + 1) {
        double cc = d[j - 1], ss = d[j];
        if (ss == 0.0) {
          if (cc != 0.0) {
            d[j] = 0.0;
#pragma unroll
            for (int k = 0; k < 4; k++) J[j * 4 + k] = -J[j * 4 + k];
          }
        } else if (cc == 0.0) {
          const bool neg = ss < 0.0;
          d[j - 1] = fabs(ss); d[j] = 0.0;
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const double t1 = J[(j - 1) * 4 + k], t2 = J[j * 4 + k];
            const double a = neg ? -t2 : t2;
            const double s1 = t1 + a;
            J[(j - 1) * 4 + k] = a;
            J[j * 4 + k] = (neg ? -s1 : s1) - t2;
          }
        } else {
          const double h = hyp(cc, ss);
          const double hy = __drcp_rn(h);
          d[j] = 0.0;
          ss = div_rcp(ss, h, hy); cc = div_rcp(cc, h, hy);
          if (cc < 0.0) { cc = -cc; ss = -ss; d[j - 1] = -h; } else d[j - 1] = h;
          const double xny = div_full(ss, 1.0 + cc);
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const double t1 = J[(j - 1) * 4 + k], t2 = J[j * 4 + k];
            const double a = t1 * cc + t2 * ss;
            J[(j - 1) * 4 + k] = a;
            J[j * 4 + k] = xny * (t1 + a) - t2;
          }
        }
      }
    // the pending constraint takes slot iq of the working set (the reference keeps it there from its selection on; no
    // slot is read by a run-time index here: a chain of selects over u[0..] would be turned into exactly that)
    S(SM_U + iq) = un; setA(iq, pend);
    iq++;
    // (selects, not `if (c == iq - 1)`: inside an equality branch the compiler rewrites the constant index as the run-time
    //  one, which sends the whole solver state to local memory)
    double dq = 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++) if (i < iq) { Rs(i, iq - 1) = d[i]; dq = d[i]; }      // the last one written is d[iq - 1]
    if (fabs(dq) <= EPS * R_norm) return false;
    R_norm = fmax(R_norm, fabs(dq));
    S(SM_RINV + iq - 1) = __drcp_rn(dq);
    return true;
  }

  // cpp:95-170.  R lives packed (upper triangle) in shared memory: the sub-diagonal entry a shifted column brings along
  // (the old diagonal) is held in `sub` until its rotation annihilates it.
  __device__ __forceinline__ bool delete_constraint(int l) {
    int qq = -1;
#pragma unroll
    for (int i = 3; i >= 1; i--) if (i < iq && getA(i) == l) qq = i;      // lowest matching slot, as the forward scan with break
    if (qq < 0) return false;
    double sub[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 1; i < 3; i++)
      if (i >= qq && i < iq - 1) {
        setA(i, getA(i + 1)); S(SM_U + i) = S(SM_U + i + 1);
#pragma unroll
        for (int k = 0; k <= i; k++) Rs(k, i) = Rs(k, i + 1);
        sub[i] = Rs(i + 1, i + 1);
      }
    // (the reference also moves slot iq -- the pending constraint -- down one slot: it lives in `un` / the caller's ip here;
    //  its zeroing of the vacated last column is not needed: the column is rewritten in full by the next add)
    iq--;
    if (iq == 0) return true;
#pragma unroll
    for (int j = 1; j < 3; j++)
      if (j >= qq && j < iq) {
        double cc = Rs(j, j), ss = sub[j];
        const double h = hyp(cc, ss);
        if (h != 0.0) {
          const double hy = __drcp_rn(h);
          cc = div_rcp(cc, h, hy); ss = div_rcp(ss, h, hy);
          if (cc < 0.0) { Rs(j, j) = -h; cc = -cc; ss = -ss; } else Rs(j, j) = h;
          const double xny = div_full(ss, 1.0 + cc);
#pragma unroll
          for (int k = j + 1; k < 4; k++)
            if (k < iq) {
              const double t1 = Rs(j, k), t2 = Rs(j + 1, k);
              const double a = t1 * cc + t2 * ss;
              Rs(j, k) = a;
              Rs(j + 1, k) = xny * (t1 + a) - t2;
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
    // reciprocals of the diagonal entries the shift / rotations changed
#pragma unroll
    for (int i = 1; i < 4; i++) if (i >= qq && i < iq) S(SM_RINV + i) = __drcp_rn(Rs(i, i));
    return true;
  }

  // Steps 1 and 2 in one pass over the 24 rows (cpp:282-342).  first = true: step 1 (psi, reset of ss / ip), then step 2;
  // first = false: the step-2 re-scan after a degenerate add (ss keeps its value, cpp:461).  The slacks are recomputed
  // from x (after a degenerate restore x is the x they were computed from: the same values).
  __device__ __forceinline__ double scan(bool first, bool vel_rows, double c0a, double c0b, double c0c, double& ss, int& ip, bool& found) {
    double psi = 0.0;
    found = false;
    if (first) { ss = 0.0; ip = 0; }
    const unsigned blocked = inA | excl;
#define ROWTEST(i, sv) do { const double s_ = (sv); if (s_ < 0.0) psi += s_; \
      if (s_ < ss && !((blocked >> (i)) & 1u)) { ss = s_; ip = (i); found = true; } } while (0)
    ROWTEST(0, -x[2] + S(SM_BB + 0)); ROWTEST(1, x[2] + S(SM_BB + 1));
    ROWTEST(2, -x[3] + S(SM_BB + 2)); ROWTEST(3, x[3] + S(SM_BB + 3));
    ROWTEST(4, -x[0] + S(SM_BB + 4)); ROWTEST(5, x[0] + S(SM_BB + 5));
    ROWTEST(6, -x[1] + S(SM_BB + 6)); ROWTEST(7, x[1] + S(SM_BB + 7));
    if (vel_rows) {
      ROWTEST(8, -x[0] + S(SM_BB + 8)); ROWTEST(9, x[0] + S(SM_BB + 9));
      ROWTEST(10, -x[1] + S(SM_BB + 10)); ROWTEST(11, x[1] + S(SM_BB + 11));
    }
#pragma unroll
    for (int g = 0; g < 6; g++) {
      const double c0 = (g < 2) ? c0a : (g < 4 ? c0b : c0c);
      const double p02 = c0 * x[g & 1] + S(SM_C2 + g) * x[2];
      ROWTEST(12 + 2 * g, (-p02 + S(SM_M3A + g) * x[3]) + S(SM_BB + 12 + 2 * g));
      ROWTEST(13 + 2 * g, (p02 + S(SM_M3B + g) * x[3]) + S(SM_BB + 13 + 2 * g));
    }
#undef ROWTEST
    return psi;
  }
  // column ip of CI (structural zeros are the -0.0 the reference's `0.0 * (-1)` leaves)
  __device__ __forceinline__ void load_column(int i, double c0a, double c0b, double c0c) {
#pragma unroll
    for (int k = 0; k < 4; k++) np[k] = -0.0;
    if (i < 12) {
      const int var = (i < 2) ? 2 : (i < 4) ? 3 : ((i & 2) ? 1 : 0);
      const double cf = (i & 1) ? 1.0 : -1.0;
#pragma unroll
      for (int k = 0; k < 4; k++) np[k] = (k == var) ? cf : np[k];
    } else {
      const int g = (i - 12) >> 1;
      const bool odd = i & 1;
      const double c0 = (g < 2) ? c0a : (g < 4 ? c0b : c0c);
      const double c2 = S(SM_C2 + g), m3 = odd ? S(SM_M3B + g) : S(SM_M3A + g);
      const double v0 = odd ? c0 : -c0;
      if (g & 1) np[1] = v0; else np[0] = v0;
      np[2] = odd ? c2 : -c2;
      np[3] = m3;
    }
  }

  // One solve.  Jf: J = L^-T of the tick's Hessian (structure: J00, J11, J22, J23, J33), Lf: L00 L11 L22 L32 L33 and
  // the reciprocals of the pivots, g0 / CE / ce0 of this SQP iteration.  xout: the solution (increment).
  __device__ __forceinline__ int solve(const double (&Lf)[5], const double (&Ly)[4], const double J23, const double tol, const double* g0,
                                       const double ce2, const double ce3, const double ce0, bool vel_rows,
                                       double c0a, double c0b, double c0c, int cap) {
    const double inf = CUDART_INF;
    it_outer = it_add = it_drop = it_degen = 0; iq = 0; inA = 0u; excl = 0u;
    Ap = 0x01010101u;               // A[i] = 0
    un = 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++) r[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 16; i++) J[i] = 0.0;
    J[0] = Ly[0]; J[5] = Ly[1]; J[10] = Ly[2]; J[14] = J23; J[15] = Ly[3];     // J(2,3) = J[3*4 + 2]
    R_norm = 1.0;
    // x = -G^-1 g0 (cpp:227-230) through the structured factor: the dense substitutions add exact zeros elsewhere
    {
      double y0 = div_rcp(g0[0], Lf[0], Ly[0]);
      double y1 = div_rcp(g0[1], Lf[1], Ly[1]);
      double y2 = div_rcp(g0[2], Lf[2], Ly[2]);
      double y3 = g0[3] - y2 * Lf[3];
      y3 = div_rcp(y3, Lf[4], Ly[3]);
      y3 = div_rcp(y3, Lf[4], Ly[3]);
      y2 = div_rcp(y2 - Lf[3] * y3, Lf[2], Ly[2]);
      y1 = div_rcp(y1, Lf[1], Ly[1]);
      y0 = div_rcp(y0, Lf[0], Ly[0]);
      x[0] = -y0; x[1] = -y1; x[2] = -y2; x[3] = -y3;
    }
    f_value = 0.5 * dot4(g0, x);
    int status = 0;
    // the equality constraint (cpp:236-276); CE = (-0, -0, ce2, ce3)
    {
      const bool allzero = (fabs(ce2) <= 1e-12) && (fabs(ce3) <= 1e-12);
      if (!allzero) {
        np[0] = -0.0; np[1] = -0.0; np[2] = ce2; np[3] = ce3;
        compute_d(); update_z();
        double t2 = 0.0;
        const double zn = dot4(z, np);
        if (fabs(dot4(z, z)) > EPS) t2 = div_full(-dot4(np, x) - ce0, zn);
#pragma unroll
        for (int k = 0; k < 4; k++) x[k] += t2 * z[k];
        un = t2;
        f_value += 0.5 * (t2 * t2) * zn;
        if (!add_constraint(-1)) return 5;
      }
    }
    enum { PH_L1, PH_L2, PH_L2A };
    int phase = PH_L1, ip = 0, l = 0, passes = 0;
    double ss = 0.0, s_ip = 0.0;
    for (;;) {
      if (phase != PH_L2A) {
        if (phase == PH_L1) {
          it_outer++;
#pragma unroll
          for (int i = 1; i < 4; i++) if (i < iq) inA |= 1u << getA(i);
          excl = 0u;
        }
        bool found;
        const double psi = scan(phase == PH_L1, vel_rows, c0a, c0b, c0c, ss, ip, found);
        if (phase == PH_L1) {
          if (fabs(psi) <= tol) break;
#pragma unroll
          for (int i = 0; i < 4; i++) { S(SM_UOLD + i) = S(SM_U + i); S(SM_XOLD + i) = x[i]; }
          S(SM_AOLD) = (double)Ap;
        }
        if (ss >= 0.0) break;
        if (found) s_ip = ss;       // none found in a re-scan: ss / ip / s[ip] keep their values, as in the reference
        load_column(ip, c0a, c0b, c0c);
        un = 0.0;
        phase = PH_L2A;
      }
      if (++passes > cap) { status = 3; break; }
      compute_d(); update_z(); update_r();
      l = 0;
      double t1 = inf, t2;
#pragma unroll
      for (int k = 1; k < 4; k++)
        if (k < iq && r[k] > 0.0) { const double tmp = div_full(S(SM_U + k), r[k]); if (tmp < t1) { t1 = tmp; l = getA(k); } }
      const double zn = dot4(z, np);
      if (fabs(dot4(z, z)) > EPS) t2 = div_full(-s_ip, zn); else t2 = inf;
      const double t = fmin(t1, t2);
      if (t >= inf) { status = 2; f_value = inf; break; }
      const double uiq = un;
      if (t2 >= inf) {
#pragma unroll
        for (int k = 0; k < 4; k++) if (k < iq) S(SM_U + k) -= t * r[k];
        un += t;
        inA &= ~(1u << l);
        if (!delete_constraint(l)) { status = 3; break; }
        it_drop++;
        continue;
      }
#pragma unroll
      for (int k = 0; k < 4; k++) x[k] += t * z[k];
      f_value += t * zn * (0.5 * t + uiq);
#pragma unroll
      for (int k = 0; k < 4; k++) if (k < iq) S(SM_U + k) -= t * r[k];
      un += t;
      if (t == t2) {
        if (!add_constraint(ip)) {
          it_degen++;
          excl |= 1u << ip;
          if (!delete_constraint(ip)) { status = 3; break; }
          inA = 0u;
          Ap = (unsigned)S(SM_AOLD);      // slots >= iq are not read before they are rewritten
#pragma unroll
          for (int i = 0; i < 4; i++)
            if (i < iq) { const int a = getA(i); if (a >= 0) inA |= 1u << a; S(SM_U + i) = S(SM_UOLD + i); }
#pragma unroll
          for (int k = 0; k < 4; k++) x[k] = S(SM_XOLD + k);
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
      s_ip = dot4(np, x) + S(SM_BB + ip);     // s[ip] at the new x: the column is still in np
    }
    if (status == 0) {
#pragma unroll
      for (int i = 0; i < 4; i++) if (x[i] != x[i]) status = 4;
    }
    return status;
  }
};
}  // namespace

#ifndef GO1_SQP_MINB
#define GO1_SQP_MINB 2
#endif

// ---------------------------------------------------------------------------------------------- launch 1: the SQP
__global__ void __launch_bounds__(SQP_THREADS, GO1_SQP_MINB) step_sqp_kernel(StepKParams P) {
  extern __shared__ __align__(16) unsigned char sqp_smem_raw[];
  const int b = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (b >= P.B) return;
  double* sm = reinterpret_cast<double*>(sqp_smem_raw) + threadIdx.x;
  const size_t B = (size_t)P.B;
  const double* S = P.state + b;
#define ST(f) S[(size_t)(f) * B]
  const StepCfgDev& c = P.cfg;
  const double dt = c.dt, Wn = c.Wn;
  const int i = P.tick[b];
  if (i < 1) return;          // no tick for this planner (the reference's i starts at 1): nothing is read or written
  // :702-704 Indexfind((i+1) dt, xyz0 = -1): first entry the time has not passed
  int j = NS;
#pragma unroll
  for (int k = NS - 1; k >= 0; k--) { const double txk = ST(S_TX + k); if (!((i + 1) * dt > txk + 0.0001)) j = k; }
  int p = (j - 1) + 1;
  const bool valid = (p >= 1 && p <= NS);
  if (!valid) p = 1;   // table overrun (UB in the reference): flagged in diag, nothing is written
  const double px = ST(S_FX + p - 1), py = ST(S_FY + p - 1);
  const double tx_p1 = ST(S_TX + p - 1), ts_p1 = ST(S_TS + p - 1);
  const int ki = (int)round(tx_p1 / dt);
  const int k_yu = i - ki;
  const double Tk = ts_p1 - k_yu * dt;
  const double Lxx_refx = ST(S_LXX + p - 1), Lyy_refy = ST(S_LYY + p - 1);
  const double tr1_ref = cosh(Wn * Tk), tr2_ref = sinh(Wn * Tk);
  double v[4];
  if (i == 1) { v[0] = Lxx_refx; v[1] = Lyy_refy; v[2] = tr1_ref; v[3] = tr2_ref; }
  else {
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = ST(S_VARI + k);
  }
  // remaining-time bounds (:745-757): functions of k_yu only -- from the table the host libm filled (the values the
  // CPU reference uses), evaluated here only outside the table
  double tr1_min, tr2_min, tr1_max, tr2_max;
  if (k_yu >= 0 && k_yu < STEP_TRTAB_ROWS) {
    const double2* tt = reinterpret_cast<const double2*>(P.trtab) + 2 * k_yu;
    const double2 qa = __ldg(tt), qb = __ldg(tt + 1);
    tr1_min = qa.x; tr2_min = qa.y; tr1_max = qb.x; tr2_max = qb.y;
  } else {
    if ((c.t_min - k_yu * dt) >= 0.001) { tr1_min = cosh(Wn * (c.t_min - k_yu * dt)); tr2_min = sinh(Wn * (c.t_min - k_yu * dt)); }
    else { tr1_min = cosh(Wn * (0.001)); tr2_min = sinh(Wn * (0.001)); }
    tr1_max = cosh(Wn * (c.t_max - k_yu * dt)); tr2_max = sinh(Wn * (c.t_max - k_yu * dt));
  }
  const double comx_f = ST(S_FEED + 0), comvx_f = ST(S_FEED + 1), comy_f = ST(S_FEED + 3), comvy_f = ST(S_FEED + 4);
  double endx = ST(S_END + 0), endy = ST(S_END + 1);
  if (i == 1) {
    const double isx = comx_f - px, esx = v[0] * 0.5, visx = (esx - isx * v[2]) / (1 / Wn * v[3]);
    const double isy = comy_f - py, esy = v[1] * 0.5, visy = (esy - isy * v[2]) / (1 / Wn * v[3]);
    endx = Wn * isx * v[3] + visx * v[2];
    endy = Wn * isy * v[3] + visy * v[2];
  }
  // objective (:1144-1173)
  const double AxO = comx_f - px, BxO = comvx_f / Wn, Cx = -0.5 * Lxx_refx;
  const double Axv = Wn * BxO, Bxv = Wn * AxO, Cxv = -endx;
  const double AyO = comy_f - py, ByO = comvy_f / Wn, Cy = -0.5 * Lyy_refy;
  const double Ayv = Wn * ByO, Byv = Wn * AyO, Cyv = -endy;
  const double aax = c.aax, aay = c.aay, aaxv = c.aaxv, aayv = c.aayv;
  // SQ = sym(SQ0): the (0,0), (1,1) entries and the 2x2 block (2..3, 2..3); everything else is an exact zero
  const double SQ00 = (0.5 * c.bbx + 0.5 * c.bbx) / 2.0, SQ11 = (0.5 * c.bby + 0.5 * c.bby) / 2.0;
  const double q22 = 0.5 * (c.rr1 + aax * AxO * AxO + aay * AyO * AyO + aaxv * Axv * Axv + aayv * Ayv * Ayv);
  const double q23 = 0.5 * (aax * AxO * BxO + aay * AyO * ByO + aaxv * Axv * Bxv + aayv * Ayv * Byv);
  const double q32 = 0.5 * (aax * BxO * AxO + aay * ByO * AyO + aaxv * Bxv * Axv + aayv * Byv * Ayv);
  const double q33 = 0.5 * (c.rr2 + aax * BxO * BxO + aay * ByO * ByO + aaxv * Bxv * Bxv + aayv * Byv * Byv);
  const double SQ22 = (q22 + q22) / 2.0, SQ23 = (q23 + q32) / 2.0, SQ32 = (q32 + q23) / 2.0, SQ33 = (q33 + q33) / 2.0;
  double Sq[4];
  Sq[0] = -c.bbx * Lxx_refx;
  Sq[1] = -c.bby * Lyy_refy;
  Sq[2] = -c.rr1 * tr1_ref + aax * AxO * Cx + aay * AyO * Cy + aaxv * Axv * Cxv + aayv * Ayv * Cyv;
  Sq[3] = -c.rr2 * tr2_ref + aax * BxO * Cx + aay * ByO * Cy + aaxv * Bxv * Cxv + aayv * Byv * Cyv;
  // G = 2 SQ (column-major G[k*4 + r] = 2 SQ[r][k]) and its factorisation, once per tick (EiQuadProg.cpp:502-515):
  // unblocked left-looking LLT, J = L^-T, c1 = trace(G), c2 = trace(J).  pd: 0 = positive definite.
  const double G00 = 2 * SQ00, G11 = 2 * SQ11, G22 = 2 * SQ22, G32 = 2 * SQ32, G23 = 2 * SQ23, G33 = 2 * SQ33;
  double Lf[5] = {0, 0, 0, 0, 0}, Ly[4] = {0, 0, 0, 0}, J23 = 0.0, tol = 0.0;
  bool not_pd = false;
  {
    const double c1 = (((0.0 + G00) + G11) + G22) + G33;
    if (G00 <= 0.0 || G11 <= 0.0 || G22 <= 0.0) not_pd = true;
    else {
      Lf[0] = sqrt(G00); Lf[1] = sqrt(G11); Lf[2] = sqrt(G22);
      Ly[0] = __drcp_rn(Lf[0]); Ly[1] = __drcp_rn(Lf[1]); Ly[2] = __drcp_rn(Lf[2]);
      Lf[3] = div_rcp(G32, Lf[2], Ly[2]);                   // L(3,2) = G(3,2) / L22, G(3,2) = 2 SQ[3][2]
      const double x33 = G33 - Lf[3] * Lf[3];
      if (x33 <= 0.0) not_pd = true;
      else {
        Lf[4] = sqrt(x33);
        Ly[3] = __drcp_rn(Lf[4]);
        // J = L^-T: J00 = 1/L00, J11 = 1/L11, J22 = 1/L22, J33 = 1/L33, J(2,3) = (0 - L32 J33) / L22
        J23 = div_rcp(-(Lf[3] * Ly[3]), Lf[2], Ly[2]);
        const double c2 = (((0.0 + Ly[0]) + Ly[1]) + Ly[2]) + Ly[3];
        tol = 24 * EPS * c1 * c2 * 100.0;
      }
    }
  }
  (void)G23;
  // lateral reachability (:1216-1246)
  double footy_max, footy_min;
  {
    const double ts_1 = ST(S_TS + 1);
    const bool wide = (i >= (round(2 * ts_1 / dt)) + 1);
    const double HW = c.half_hip_width, FW = c.foot_width;
    if (p % 2 == 0) { footy_min = -(2 * HW + 0.03); footy_max = wide ? -(FW + 0.01) : -(HW - 0.03); }
    else { footy_max = 2 * HW + 0.03; footy_min = wide ? FW + 0.01 : HW - 0.03; }
  }
  // coefficient rows that depend on the state only (:1341-1411); the six +- pairs of rows 12..23
  const double CCx = comx_f - px, CCy = comy_f - py;
  const double sh_dt = c.sh_dt, ch_dt = c.ch_dt;      // sinh / cosh(Wn dt): instance-independent, from the host
  const double AA = Wn * sh_dt;
  const double c0a = AA * Wn, c0b = ch_dt * Wn, c0c = Wn;          // AA1x = AA1y, VAA1x = VAA1y, VAA1x1 = VAA1y1
  double nd3a4, nd3b4, nd3a5, nd3b5;      // -(d3) of rows 20..23 (their right-hand sides use a dt / 2, the matrix a dt)
  {
    const double BBx = (Wn * Wn) * CCx * ch_dt, BBy = (Wn * Wn) * CCy * ch_dt;
    const double AA2x = -2 * AA * CCx * Wn, AA3x = 2 * BBx;
    const double AA2y = -2 * AA * CCy * Wn, AA3y = 2 * BBy;
    const double VAA = ch_dt;
    const double VBBx = Wn * CCx * sh_dt, VBBy = Wn * CCy * sh_dt;
    const double VAA2x = -2 * VAA * CCx * Wn, VAA3x = 2 * VBBx - 2 * comvx_f;
    const double VAA2y = -2 * VAA * CCy * Wn, VAA3y = 2 * VBBy - 2 * comvy_f;
    const double VAA2x1 = -2 * CCx * Wn, VAA3x1 = -2 * comvx_f;
    const double VAA2y1 = -2 * CCy * Wn, VAA3y1 = -2 * comvy_f;
    sm[(SM_C2 + 0) * SQP_THREADS] = AA2x; sm[(SM_C2 + 1) * SQP_THREADS] = AA2y;
    sm[(SM_C2 + 2) * SQP_THREADS] = VAA2x; sm[(SM_C2 + 3) * SQP_THREADS] = VAA2y;
    sm[(SM_C2 + 4) * SQP_THREADS] = VAA2x1; sm[(SM_C2 + 5) * SQP_THREADS] = VAA2y1;
    // CI(3, row) = c3 * (-1); even row: c3 = X - 2 a_max (dt), odd row: c3 = -(X - 2 a_min (dt))
    sm[(SM_M3A + 0) * SQP_THREADS] = (AA3x - 2 * c.comax_max) * (-1); sm[(SM_M3B + 0) * SQP_THREADS] = (-(AA3x - 2 * c.comax_min)) * (-1);
    sm[(SM_M3A + 1) * SQP_THREADS] = (AA3y - 2 * c.comay_max) * (-1); sm[(SM_M3B + 1) * SQP_THREADS] = (-(AA3y - 2 * c.comay_min)) * (-1);
    sm[(SM_M3A + 2) * SQP_THREADS] = (VAA3x - 2 * c.comax_max * dt) * (-1); sm[(SM_M3B + 2) * SQP_THREADS] = (-(VAA3x - 2 * c.comax_min * dt)) * (-1);
    sm[(SM_M3A + 3) * SQP_THREADS] = (VAA3y - 2 * c.comay_max * dt) * (-1); sm[(SM_M3B + 3) * SQP_THREADS] = (-(VAA3y - 2 * c.comay_min * dt)) * (-1);
    sm[(SM_M3A + 4) * SQP_THREADS] = (VAA3x1 - 2 * c.comax_max * dt) * (-1); sm[(SM_M3B + 4) * SQP_THREADS] = (-(VAA3x1 - 2 * c.comax_min * dt)) * (-1);
    sm[(SM_M3A + 5) * SQP_THREADS] = (VAA3y1 - 2 * c.comay_max * dt) * (-1); sm[(SM_M3B + 5) * SQP_THREADS] = (-(VAA3y1 - 2 * c.comay_min * dt)) * (-1);
    nd3a4 = -(VAA3x1 - 2 * c.comax_max * dt / 2.0); nd3b4 = -(-(VAA3x1 - 2 * c.comax_min * dt / 2.0));
    nd3a5 = -(VAA3y1 - 2 * c.comay_max * dt / 2.0); nd3b5 = -(-(VAA3y1 - 2 * c.comay_min * dt / 2.0));
  }
  const bool vel_rows = (k_yu != 0);
  const double fxv_max_dt = c.footx_vmax * dt, fxv_min_dt = c.footx_vmin * dt, fyv_max_dt = c.footy_vmax * dt, fyv_min_dt = c.footy_vmin * dt;

  int* DG = P.diag ? P.diag + b : nullptr;
#define DGW(f, val) do { if (DG) DG[(size_t)(f) * B] = (val); } while (0)
  int n_solved = 0;
  const bool do_solve = (Tk >= 0.1 * ts_p1);
#pragma unroll 1
  for (int it = 1; it <= P.n_sqp; it++) {
    if (!do_solve) { v[0] = Lxx_refx; v[1] = Lyy_refy; v[2] = tr1_ref; v[3] = tr2_ref; continue; }
    // g0 = G v + Sq in the dense order (:1164-1165): the structural zeros of G add exact zeros
    double g0[4];
    g0[0] = (2 * SQ00) * v[0] + Sq[0];
    g0[1] = (2 * SQ11) * v[1] + Sq[1];
    g0[2] = ((2 * SQ22) * v[2] + (2 * SQ23) * v[3]) + Sq[2];
    g0[3] = ((2 * SQ32) * v[2] + (2 * SQ33) * v[3]) + Sq[3];
    // linearised tr1^2 - tr2^2 = 1 (:1188-1190): CE = -(2 [0 0 tr1 -tr2])', ce0 = 1 - (tr1^2 - tr2^2)
    double q = 0.0;
    q += v[2] * v[2];
    q += (v[3] * (-1)) * v[3];
    const double ce0 = -q + 1;
    const double ce2 = (2 * v[2]) * (-1), ce3 = ((-2) * v[3]) * (-1);
    // right-hand sides b = bound - A v (:1193-1455)
#define BBW(r, val) sm[(SM_BB + (r)) * SQP_THREADS] = (val)
    BBW(0, -(v[2]) + tr1_max);
    BBW(1, -((-1.0) * v[2]) - tr1_min);
    BBW(2, -(v[3]) + tr2_max);
    BBW(3, -((-1.0) * v[3]) - tr2_min);
    BBW(4, -(v[0]) + c.footx_max);
    BBW(5, -((-1.0) * v[0]) - c.footx_min);
    BBW(6, -(v[1]) + footy_max);
    BBW(7, -((-1.0) * v[1]) - footy_min);
    if (vel_rows) {
      BBW(8, -(v[0] - Lxx_refx - fxv_max_dt));
      BBW(9, v[0] - Lxx_refx - fxv_min_dt);
      BBW(10, -(v[1] - Lyy_refy - fyv_max_dt));
      BBW(11, v[1] - Lyy_refy - fyv_min_dt);
    } else {
      BBW(8, 0.0); BBW(9, 0.0); BBW(10, 0.0); BBW(11, 0.0);
    }
#pragma unroll
    for (int g = 0; g < 6; g++) {
      const double c0 = (g < 2) ? c0a : (g < 4 ? c0b : c0c);
      const double c2 = sm[(SM_C2 + g) * SQP_THREADS];
      const double p0 = c0 * v[g & 1], p2 = c2 * v[2];
      // even row: A = (c0, c2, c3a): b = ((-c0) v + (-c2) tr1) + (-d3a) tr2;  odd row: A = (-c0, -c2, c3b)
      double nda = sm[(SM_M3A + g) * SQP_THREADS], ndb = sm[(SM_M3B + g) * SQP_THREADS];
      if (g == 4) { nda = nd3a4; ndb = nd3b4; }
      if (g == 5) { nda = nd3a5; ndb = nd3b5; }
      BBW(12 + 2 * g, (-p0 + -p2) + nda * v[3]);
      BBW(13 + 2 * g, (p0 + p2) + ndb * v[3]);
    }
#undef BBW
    int st;
    double X[4];
    Qp4 qp;
    qp.sm = sm;
    if (not_pd) {
      st = 1;
#pragma unroll
      for (int k = 0; k < 4; k++) X[k] = v[k];
      qp.iq = 0; qp.it_outer = qp.it_add = qp.it_drop = qp.it_degen = 0;
      qp.Ap = 0x01010101u;
    } else {
      st = qp.solve(Lf, Ly, J23, tol, g0, ce2, ce3, ce0, vel_rows, c0a, c0b, c0c, P.cap);
#pragma unroll
      for (int k = 0; k < 4; k++) X[k] = qp.x[k];
    }
    if (n_solved < STEP_MAX_SQP) {
      const int o = STEP_DIAG_HEAD + n_solved * STEP_DIAG_PER;
      DGW(o + 0, st); DGW(o + 1, st == 1 ? 0 : qp.iq);
      DGW(o + 2, qp.it_outer); DGW(o + 3, qp.it_add); DGW(o + 4, qp.it_drop); DGW(o + 5, qp.it_degen);
#pragma unroll
      for (int k = 0; k < 4; k++) DGW(o + 6 + k, (st != 1 && k < qp.iq) ? qp.getA(k) : -99);
      DGW(o + 10, -99);                  // the working set never holds more than n = 4 constraints
    }
    n_solved++;
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] += X[k];   // :795-798, whatever the status
  }
  for (int qd = n_solved; qd < STEP_MAX_SQP; qd++) {      // every diag entry is defined: unused slots read -1, 0...
    DGW(STEP_DIAG_HEAD + qd * STEP_DIAG_PER, -1);
    for (int k = 1; k < STEP_DIAG_PER; k++) DGW(STEP_DIAG_HEAD + qd * STEP_DIAG_PER + k, 0);
  }
  DGW(0, valid ? p : -1); DGW(1, k_yu); DGW(4, n_solved);
  // hand-over to the next two kernels through rows of the output buffer (36, 37 are final; 0, 1, 28, 29 scratch)
  double* O = P.out + b;
  O[(size_t)36 * B] = v[0]; O[(size_t)37 * B] = v[1];
  O[(size_t)0 * B] = v[2]; O[(size_t)1 * B] = v[3];
  O[(size_t)28 * B] = valid ? (double)p : 0.0;
  O[(size_t)29 * B] = (double)k_yu;
#undef DGW
#undef ST
}

// ---------------------------------------------------------------------------------------------- launch 2: back-end
// The three columns of the 7x7 inverse CoM_height_solve uses (its right-hand side is zero outside entries 2..4), by the
// reference's row-pivoted Gauss-Jordan with the matrix in registers: rinv[i*3 + c] = A^-1(i, 2 + c).
// Pivot step K and candidate row I are TEMPLATE constants: the row swap sits behind `if (piv == I)` (the planners of
// a warp nearly always agree on the pivot, so one swap runs instead of six predicated ones), and an equality branch on a
// loop variable would be rewritten by the compiler as a run-time index, i.e. local memory.
// The row normalisation divides by the pivot through ONE reciprocal per pivot (div_rcp: bit-identical quotients).
template <int K, int I>
struct GjSwap {
  static __device__ __forceinline__ void run(double (&a)[49], double (&rinv)[21], int piv) {
    if (piv == I) {
#pragma unroll
      for (int j = K; j < 7; j++) { const double t = a[K * 7 + j]; a[K * 7 + j] = a[I * 7 + j]; a[I * 7 + j] = t; }
#pragma unroll
      for (int c = 0; c < 3; c++) { const double t = rinv[K * 3 + c]; rinv[K * 3 + c] = rinv[I * 3 + c]; rinv[I * 3 + c] = t; }
    }
    GjSwap<K, I + 1>::run(a, rinv, piv);
  }
};
template <int K>
struct GjSwap<K, 7> {
  static __device__ __forceinline__ void run(double (&)[49], double (&)[21], int) {}
};
template <int K>
struct GjStep {
  static __device__ __forceinline__ void run(double (&a)[49], double (&rinv)[21]) {
    constexpr int n = 7;
    int piv = K;
    double best = fabs(a[K * n + K]);
#pragma unroll
    for (int i = K + 1; i < n; i++) { const double v = fabs(a[i * n + K]); const bool gt = v > best; best = gt ? v : best; piv = gt ? i : piv; }
    GjSwap<K, K + 1>::run(a, rinv, piv);
    const double d = a[K * n + K];
    const double y = __drcp_rn(d);
    const bool ok = fabs(d) > 1e-290 && fabs(d) < 1e290;      // outside: plain IEEE divisions (singular / overflowing pivot)
#pragma unroll
    for (int j = K; j < n; j++) a[K * n + j] = ok ? div_rcp(a[K * n + j], d, y) : a[K * n + j] / d;
#pragma unroll
    for (int c = 0; c < 3; c++) rinv[K * 3 + c] = ok ? div_rcp(rinv[K * 3 + c], d, y) : rinv[K * 3 + c] / d;
#pragma unroll
    for (int i = 0; i < n; i++) {
      if (i == K) continue;
      const double f = a[i * n + K];
#pragma unroll
      for (int j = K; j < n; j++) a[i * n + j] = __dsub_rn(a[i * n + j], __dmul_rn(f, a[K * n + j]));
#pragma unroll
      for (int c = 0; c < 3; c++) rinv[i * 3 + c] = __dsub_rn(rinv[i * 3 + c], __dmul_rn(f, rinv[K * 3 + c]));
    }
    GjStep<K + 1>::run(a, rinv);
  }
};
template <>
struct GjStep<7> {
  static __device__ __forceinline__ void run(double (&)[49], double (&)[21]) {}
};
__device__ __forceinline__ void gj_inverse7_cols234(double (&a)[49], double (&rinv)[21]) {
#pragma unroll
  for (int i = 0; i < 7; i++)
#pragma unroll
    for (int c = 0; c < 3; c++) rinv[i * 3 + c] = (i == 2 + c) ? 1.0 : 0.0;
  GjStep<0>::run(a, rinv);
}

// NLPClass::CoM_height_solve (NLPClass_sqp.cpp:2361-2473) for the samples i, i+1, i+2
__device__ __forceinline__ void com_height_solve(int i, int bjx1, double ts1, double tx1, double f0, double f1, double hcom, double dt,
                                                 double comz[3], double comvz[3], double comaz[3], double* co_out, size_t co_stride) {
  if (co_out) {        // constant polynomial z = hcom unless the fit below replaces it
#pragma unroll
    for (int r = 0; r < 6; r++) co_out[(size_t)r * co_stride] = 0.0;
    co_out[(size_t)6 * co_stride] = hcom; co_out[(size_t)7 * co_stride] = 0.0;
  }
  if (bjx1 >= 2) {
    const double tp[3] = {0.0001, ts1 / 2 + 0.0001, ts1 + 0.0001};
    double A[49], Rinv[21];
#pragma unroll
    for (int g = 0; g < 3; g++) {
      double pw[7];
      powi_all(tp[g], pw);
      const int r_pos = (g == 0) ? 2 : (g == 1 ? 3 : 4), r_vel = (g == 0) ? 0 : 5, r_acc = (g == 0) ? 1 : 6;
      { double* a = A + 7 * r_pos; a[0] = pw[6]; a[1] = pw[5]; a[2] = pw[4]; a[3] = pw[3]; a[4] = pw[2]; a[5] = pw[1]; a[6] = 1; }
      if (g != 1) {
        { double* a = A + 7 * r_vel; a[0] = 6 * pw[5]; a[1] = 5 * pw[4]; a[2] = 4 * pw[3]; a[3] = 3 * pw[2]; a[4] = 2 * pw[1]; a[5] = 1; a[6] = 0; }
        { double* a = A + 7 * r_acc; a[0] = 30 * pw[4]; a[1] = 20 * pw[3]; a[2] = 12 * pw[2]; a[3] = 6 * pw[1]; a[4] = 2; a[5] = 0; a[6] = 0; }
      }
    }
    gj_inverse7_cols234(A, Rinv);
    const double plan3[3] = {f0 + hcom, (f0 + f1) / 2 + hcom, f1 + hcom};
    double co[7];
#pragma unroll
    for (int r = 0; r < 7; r++) {
      double acc = 0.0;      // the reference's sum over k = 0..6 adds exact zeros for k = 0, 1, 5, 6
#pragma unroll
      for (int c = 0; c < 3; c++) acc = __dadd_rn(acc, __dmul_rn(Rinv[3 * r + c], plan3[c]));
      co[r] = acc;
    }
    if (co_out) {
#pragma unroll
      for (int r = 0; r < 7; r++) co_out[(size_t)r * co_stride] = co[r];
      co_out[(size_t)7 * co_stride] = round(tx1 / dt);
    }
#pragma unroll
    for (int jxx = 1; jxx <= 3; jxx++) {
      const double t = (i + jxx - round(tx1 / dt)) * dt;
      double pw[7];
      powi_all(t, pw);
      const double p[7] = {pw[6], pw[5], pw[4], pw[3], pw[2], pw[1], 1};
      const double v[7] = {6 * pw[5], 5 * pw[4], 4 * pw[3], 3 * pw[2], 2 * pw[1], 1, 0};
      const double a[7] = {30 * pw[4], 20 * pw[3], 12 * pw[2], 6 * pw[1], 2, 0, 0};
      double z = 0.0, vz = 0.0, az = 0.0;
#pragma unroll
      for (int k = 0; k < 7; k++) { z = __dadd_rn(z, __dmul_rn(p[k], co[k])); vz = __dadd_rn(vz, __dmul_rn(v[k], co[k])); az = __dadd_rn(az, __dmul_rn(a[k], co[k])); }
      comz[jxx - 1] = z; comvz[jxx - 1] = vz; comaz[jxx - 1] = az;
    }
  } else {
#pragma unroll
    for (int q = 0; q < 3; q++) { comz[q] = hcom; comvz[q] = 0; comaz[q] = 0; }
  }
}

// ---------------------------------------------------------------------------------------------- launch 2: CoM height
// ts_new (:888), the two table entries CoM_height_solve reads from the UPDATED tables, the 7x7 solve.  Register-heavy
// (the elimination keeps 70 doubles in registers), few memory accesses.  Writes out rows 35 (ts_new), 2 / 5 / 8 / 23 / 26
// (final) and 9, 10 (scratch: comz at i+1, i+2 for the ZMP / DCM of the last kernel).
#ifndef GO1_POST_MINB
#define GO1_POST_MINB 2
#endif
__global__ void __launch_bounds__(128, GO1_POST_MINB) step_height_kernel(StepKParams P) {
  const int b = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (b >= P.B) return;
  const size_t B = (size_t)P.B;
  const double* S = P.state + b;
  const double* IN = P.in + b;
  double* O = P.out + b;
#define ST(f) S[(size_t)(f) * B]
#define INP(f) IN[(size_t)(f) * B]
  const StepCfgDev& c = P.cfg;
  const double dt = c.dt, Wn = c.Wn;
  const int i = P.tick[b];
  if (i < 1) return;
  const double v2 = O[(size_t)0 * B], v3 = O[(size_t)1 * B];
  const int pv = (int)O[(size_t)28 * B];
  const int k_yu = (int)O[(size_t)29 * B];
  const int p = pv >= 1 ? pv : 1;
  const int bp = (int)ST(S_BJX1);
  const int b1 = bp >= 1 && bp <= NS ? bp : 1, b2 = bp >= 2 && bp <= NS + 1 ? bp : 2;
  const double ts_new = k_yu * dt + log(v2 + v3) / Wn;
  O[(size_t)35 * B] = ts_new;
  double hz_z[3], hz_vz[3], hz_az[3];
  if (c.ext_height) {
#pragma unroll
    for (int q = 0; q < 3; q++) { hz_z[q] = INP(I_CZ + q); hz_az[q] = INP(I_CAZ + q); hz_vz[q] = 0.0; }
    hz_vz[0] = INP(I_CVZ);
  } else {
    const double fz_b2 = ST(S_FZ + b2 - 2), fz_b1 = ST(S_FZ + b1 - 1);
    // _ts(b1-1) and _tx(b1-1) after the write-back (:886-909): _ts(p-1) = ts_new, _tx(k) = _tx(k-1) + _ts(k-1) for k >= p
    const double ts_b1 = (b1 == p) ? ts_new : ST(S_TS + b1 - 1);
    double tx_b1;
    if (b1 - 1 >= 1 && b1 - 1 >= p) {
      double cur = ST(S_TX + p - 1);
      for (int k = p; k <= b1 - 1; k++) cur = cur + ((k == p) ? ts_new : ST(S_TS + k - 1));
      tx_b1 = cur;
    } else {
      tx_b1 = ST(S_TX + b1 - 1);
    }
    com_height_solve(i, bp <= NS ? bp : 0, ts_b1, tx_b1, fz_b2, fz_b1, c.hcom, dt, hz_z, hz_vz, hz_az, P.hz_co ? P.hz_co + b : nullptr, B);
  }
  O[(size_t)2 * B] = hz_z[0]; O[(size_t)5 * B] = hz_vz[0]; O[(size_t)8 * B] = hz_az[0];
  O[(size_t)23 * B] = hz_az[1]; O[(size_t)26 * B] = hz_az[2];
  O[(size_t)9 * B] = hz_z[1]; O[(size_t)10 * B] = hz_z[2];
#undef ST
#undef INP
}

// ---------------------------------------------------------------------------------------------- launch 3: back-end
// Write-back of step length / width / period and the step table, LIPM roll-out, feedback blend, integer step indices,
// outputs.  Memory-bound (about 120 coalesced loads / stores per planner), few registers: many warps per SM.
__global__ void __launch_bounds__(128, 4) step_finish_kernel(StepKParams P) {
  const int b = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (b >= P.B) return;
  const size_t B = (size_t)P.B;
  const double* S = P.state + b;        // read side (state before the tick)
  double* SO = P.state_out + b;         // write side; holds a copy of the read side when it is another buffer
  const double* IN = P.in + b;
  double* O = P.out + b;
#define ST(f) S[(size_t)(f) * B]
#define STW(f) SO[(size_t)(f) * B]
#define INP(f) IN[(size_t)(f) * B]
  const StepCfgDev& c = P.cfg;
  const double dt = c.dt, Wn = c.Wn;
  const int i = P.tick[b];
  if (i < 1) return;
  // hand-over of the two kernels before
  double v[4];
  v[0] = O[(size_t)36 * B]; v[1] = O[(size_t)37 * B]; v[2] = O[(size_t)0 * B]; v[3] = O[(size_t)1 * B];
  const int pv = (int)O[(size_t)28 * B];
  const double ts_new = O[(size_t)35 * B];
  double hz_z[3], hz_az[3];
  hz_z[0] = O[(size_t)2 * B]; hz_z[1] = O[(size_t)9 * B]; hz_z[2] = O[(size_t)10 * B];
  hz_az[0] = O[(size_t)8 * B]; hz_az[1] = O[(size_t)23 * B]; hz_az[2] = O[(size_t)26 * B];
  const bool valid = pv >= 1;
  const int p = valid ? pv : 1;
  const double px = ST(S_FX + p - 1), py = ST(S_FY + p - 1);
  const double tx_p1 = ST(S_TX + p - 1);
  const double comx_f = ST(S_FEED + 0), comy_f = ST(S_FEED + 3);
  // (:896-901)
  const double isx = comx_f - px, esx = v[0] * 0.5, visx = (esx - isx * v[2]) / (1 / Wn * v[3]);
  const double isy = comy_f - py, esy = v[1] * 0.5, visy = (esy - isy * v[2]) / (1 / Wn * v[3]);
  const double fx_next = px + v[0], fy_next = py + v[1];
  if (P.lipm) {
    double* LP = P.lipm + b;
    LP[0] = isx; LP[B] = visx; LP[2 * B] = isy; LP[3 * B] = visy; LP[4 * B] = px; LP[5 * B] = py;
  }
  // _ts(p-1) = ts_new and the running sum _tx(k) = _tx(k-1) + _ts(k-1) for k >= p (:906-909), in the reference's order;
  // in the same pass: the two index searches against the UPDATED table (:1031-1041) and the store of the new _tx
  // entries (every load of column k precedes its store: in-place safe)
  int jA = NS, jB = NS;
  {
    double cur = tx_p1;
#pragma unroll
    for (int k = 0; k < NS; k++) {
      double txk;
      if (k >= 1 && k >= p) {
        const double tsk = (k == p) ? ts_new : ST(S_TS + (k >= 1 ? k - 1 : 0));
        cur = cur + tsk;
        txk = cur;
        if (valid) STW(S_TX + k) = txk;
      } else {
        txk = ST(S_TX + k);
      }
      if (jA == NS && !(i * dt >= txk)) jA = k;
      if (jB == NS && !((i + 1) * dt >= txk)) jB = k;
    }
  }
  // LIPM roll-out of samples i, i+1, i+2 (:938-955)
  double comx[3], comy[3], comvx[3], comvy[3], comax[3], comay[3], zmpx[3], zmpy[3], dcmx[3], dcmy[3];
#pragma unroll
  for (int jxx = 1; jxx <= 3; jxx++) {
    const int q = jxx - 1;
    const double ch = c.ch_w[q], sh = c.sh_w[q];       // cosh / sinh(Wn dt jxx), from the host
    comx[q] = isx * ch + visx * 1 / Wn * sh + px;
    comy[q] = isy * ch + visy * 1 / Wn * sh + py;
    comvx[q] = Wn * isx * sh + visx * ch;
    comvy[q] = Wn * isy * sh + visy * ch;
    comax[q] = (Wn * Wn) * isx * ch + visx * Wn * sh;
    comay[q] = (Wn * Wn) * isy * ch + visy * Wn * sh;
    const double hz = (hz_z[q] - INP(I_ZSC + q)) / (hz_az[q] + c.ggg);
    zmpx[q] = comx[q] - hz * comax[q];
    zmpy[q] = comy[q] - hz * comay[q];
    dcmx[q] = comx[q] + comvx[q] * sqrt(hz);
    dcmy[q] = comy[q] + comvy[q] * sqrt(hz);
  }
  // feedback blend (:963-972, :1017-1022)
  double e0 = INP(I_EST + 0), e3 = INP(I_EST + 3);
  if (p % 2 == 0) { e0 = e0 - INP(I_LF + 0); e3 = e3 - INP(I_LF + 1); }
  else { e0 = e0 - INP(I_RF + 0); e3 = e3 - INP(I_RF + 1); }
  const double lx = c.lamda[0], lvx = c.lamda[1], ly = c.lamda[2], lvy = c.lamda[3];
  // integer step indices against the UPDATED table (:1031-1041)
  const int bjxx = jA, bjx1 = jB;
  // foot tables at the two entries the outputs read, taken BEFORE any write (in-place safe)
  const int bq0 = bjxx < NS ? bjxx : NS - 1, bq1 = bjxx + 1 < NS ? bjxx + 1 : NS - 1;
  const bool upd = valid && p < NS;
  const double fx0 = (upd && bq0 == p) ? fx_next : ST(S_FX + bq0), fx1 = (upd && bq1 == p) ? fx_next : ST(S_FX + bq1);
  const double fy0 = (upd && bq0 == p) ? fy_next : ST(S_FY + bq0), fy1 = (upd && bq1 == p) ? fy_next : ST(S_FY + bq1);
  const double fz0 = ST(S_FZ + bq0), fz1 = ST(S_FZ + bq1);
  if (valid) {
#pragma unroll
    for (int k = 0; k < 4; k++) STW(S_VARI + k) = v[k];
    STW(S_LXX + p - 1) = v[0];
    STW(S_LYY + p - 1) = v[1];
    STW(S_TS + p - 1) = ts_new;
    if (p < NS) { STW(S_FX + p) = fx_next; STW(S_FY + p) = fy_next; }
    STW(S_END + 0) = Wn * isx * v[3] + visx * v[2];
    STW(S_END + 1) = Wn * isy * v[3] + visy * v[2];
    STW(S_FEED + 0) = ((1 - lx) * (comx[0] - px) + (lx) * e0) + px;
    STW(S_FEED + 1) = (1 - lvx) * comvx[0] + (lvx) * INP(I_EST + 1);
    STW(S_FEED + 2) = (1 - lx) * comax[0] + lx * INP(I_EST + 2);
    STW(S_FEED + 3) = ((1 - ly) * (comy[0] - py) + (ly) * e3) + py;
    STW(S_FEED + 4) = (1 - lvy) * comvy[0] + (lvy) * INP(I_EST + 4);
    STW(S_FEED + 5) = (1 - ly) * comay[0] + ly * INP(I_EST + 5);
    STW(S_BJX1) = (double)bjx1;
  }
#define OUT(f, val) O[(size_t)(f) * B] = (val)
  OUT(0, comx[0]); OUT(1, comy[0]);                      // 2, 5, 8, 23, 26, 35: written by step_height_kernel
  OUT(3, comvx[0]); OUT(4, comvy[0]);
  OUT(6, comax[0]); OUT(7, comay[0]);
  OUT(9, zmpx[0]); OUT(10, zmpy[0]); OUT(11, dcmx[0]); OUT(12, dcmy[0]);
  OUT(13, zmpx[1]); OUT(14, zmpy[1]); OUT(15, dcmx[1]); OUT(16, dcmy[1]);
  OUT(17, zmpx[2]); OUT(18, zmpy[2]); OUT(19, dcmx[2]); OUT(20, dcmy[2]);
  OUT(21, comax[1]); OUT(22, comay[1]);
  OUT(24, comax[2]); OUT(25, comay[2]);
  OUT(27, (double)bjxx);
  OUT(28, fx0); OUT(29, fx1); OUT(30, fy0); OUT(31, fy1);
  OUT(32, fz0); OUT(33, fz1);
  OUT(34, (double)(p - 1));
  // rows 36, 37 (v[0], v[1]) were written by step_sqp_kernel
  if (P.diag) { int* DG = P.diag + b; DG[(size_t)2 * B] = bjxx; DG[(size_t)3 * B] = bjx1; }
#undef OUT
#undef ST
#undef STW
#undef INP
}

cudaError_t step_sqp_launch(StepKParams P, cudaStream_t st) {
  const size_t smem = (size_t)SM_SLOTS * SQP_THREADS * sizeof(double);
  static std::mutex mu;              // handles on several host threads may launch concurrently
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  {
    std::lock_guard<std::mutex> lk(mu);
    if (!attr_set[dev & 63]) {
      cudaError_t e = cudaFuncSetAttribute(step_sqp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      attr_set[dev & 63] = true;
    }
  }
  step_sqp_kernel<<<(P.B + SQP_THREADS - 1) / SQP_THREADS, SQP_THREADS, smem, st>>>(P);
  return cudaGetLastError();
}
// state_out must already hold a copy of state when it is another buffer (the caller's job: api.cu overlaps that copy
// with step_sqp_kernel on a side stream)
cudaError_t step_height_launch(StepKParams P, cudaStream_t st) {
  step_height_kernel<<<(P.B + 127) / 128, 128, 0, st>>>(P);
  return cudaGetLastError();
}
cudaError_t step_post_launch(StepKParams P, cudaStream_t st) {
  step_finish_kernel<<<(P.B + 127) / 128, 128, 0, st>>>(P);
  return cudaGetLastError();
}

}  // namespace go1
