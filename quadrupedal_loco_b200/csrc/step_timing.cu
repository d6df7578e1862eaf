// step_timing.cu -- step-location / step-timing SQP tick, latency mapping: one WARP per planner (batches below ~1000
// planners, where a thread-per-planner launch leaves the GPU empty; large batches take step_sqp.cu).
//
// Replaces, for a batch of independent planners, one 40 Hz tick of
//   NLPClass::step_timing_opti_loop      NLP/src/NLP/NLPClass_sqp.cpp:693-1102
//   step_timing_object_function          :1144-1173
//   step_timing_constraints              :1175-1458
//   solve_stepping_timing / Solve        :1613-1653   (QP: n = 4, p = 1, m = 24)
//   Indexfind                            :1105-1141
// (NLP = unitree_ros/mosek_nlp_kmp).  K SQP iterations (reference: 3), the write-back of step
// length / width / period, the LIPM roll-out of the next three samples, the feedback blend and
// the integer step indices and CoM_height_solve (:2361-2473: 6th-order vertical CoM polynomial,
// 7x7 inverse) run in ONE launch; the QP matrices live in the thread's local memory and never
// touch HBM.
//
// Layout: structure of arrays, element-major / batch-minor -- field f of instance b is at
// [f * B + b] -- so a warp's 32 instances read and write 256 contiguous bytes per field.
// Compiled with -fmad=false: same operation order as the CPU oracle, no FMA contraction (measured:
// contraction would make the thread-per-planner tick 10-20 % faster and keep the QP parity at 1e-9,
// but the step period it writes back then differs from the host's by an ulp, which the badly
// conditioned swing-foot fit downstream amplifies to 1e-6 -- fidelity wins).
#include <cuda_runtime.h>
#include "gi_warp.cuh"
#include "kernels.h"
#include "powi.cuh"

namespace go1 {

namespace {
constexpr int NS = 27;
// state fields (doubles)
constexpr int S_TS = 0, S_TX = 27, S_FX = 54, S_FY = 81, S_FZ = 108, S_LXX = 135, S_LYY = 162, S_FEED = 189, S_VARI = 195, S_END = 199, S_BJX1 = 201;
// input fields
constexpr int I_EST = 0, I_RF = 6, I_LF = 8, I_CZ = 10, I_CAZ = 13, I_ZSC = 16, I_CVZ = 19;
}  // namespace

// libm calls of the tick's front-end, out of line: eight inlined cosh / sinh expansions are ~1 k SASS instructions of a
// kernel that stalls on instruction fetch; the routines (and therefore the results) are the same
static __device__ __noinline__ double cosh_nl(double x) { return cosh(x); }
static __device__ __noinline__ double sinh_nl(double x) { return sinh(x); }

__device__ void gj_inverse7_warp(double* M, int lane);

// CoM_height_solve, warp-cooperative: the 7 x 7 inverse one lane per column, the coefficient
// vector one lane per row, the three samples one lane each; same operation order per element as
// com_height_solve.  M: >= 7*14 + 7 doubles of the warp's shared memory.
__device__ void com_height_solve_warp(int i, int bjx1, double ts1, double tx1, double f0, double f1, double hcom, double dt,
                                      double comz[3], double comvz[3], double comaz[3], double* M, int lane) {
  if (bjx1 >= 2) {
    const double tp[3] = {0.0001, ts1 / 2 + 0.0001, ts1 + 0.0001};
    const int rowt[7] = {0, 0, 0, 1, 2, 2, 2}, kind[7] = {1, 2, 0, 0, 0, 1, 2};
    __syncwarp();
    if (lane < 7) {
      const int r = lane;
      const double t = tp[rowt[r]];
      double* a = M + 14 * r;
      if (kind[r] == 0) { a[0] = powi(t, 6); a[1] = powi(t, 5); a[2] = powi(t, 4); a[3] = powi(t, 3); a[4] = powi(t, 2); a[5] = powi(t, 1); a[6] = 1; }
      else if (kind[r] == 1) { a[0] = 6 * powi(t, 5); a[1] = 5 * powi(t, 4); a[2] = 4 * powi(t, 3); a[3] = 3 * powi(t, 2); a[4] = 2 * powi(t, 1); a[5] = 1; a[6] = 0; }
      else { a[0] = 30 * powi(t, 4); a[1] = 20 * powi(t, 3); a[2] = 12 * powi(t, 2); a[3] = 6 * powi(t, 1); a[4] = 2; a[5] = 0; a[6] = 0; }
    }
    __syncwarp();
    gj_inverse7_warp(M, lane);
    const double plan[7] = {0, 0, f0 + hcom, (f0 + f1) / 2 + hcom, f1 + hcom, 0, 0};
    double* co = M + 98;
    if (lane < 7) { double acc = 0.0; for (int k = 0; k < 7; k++) acc = __dadd_rn(acc, __dmul_rn(M[14 * lane + 7 + k], plan[k])); co[lane] = acc; }
    __syncwarp();
    double z = 0.0, vz = 0.0, az = 0.0;
    if (lane < 3) {
      const int jxx = lane + 1;
      const double t = (i + jxx - round(tx1 / dt)) * dt;
      const double p[7] = {powi(t, 6), powi(t, 5), powi(t, 4), powi(t, 3), powi(t, 2), powi(t, 1), 1};
      const double v[7] = {6 * powi(t, 5), 5 * powi(t, 4), 4 * powi(t, 3), 3 * powi(t, 2), 2 * powi(t, 1), 1, 0};
      const double a[7] = {30 * powi(t, 4), 20 * powi(t, 3), 12 * powi(t, 2), 6 * powi(t, 1), 2, 0, 0};
      for (int k = 0; k < 7; k++) { z = __dadd_rn(z, __dmul_rn(p[k], co[k])); vz = __dadd_rn(vz, __dmul_rn(v[k], co[k])); az = __dadd_rn(az, __dmul_rn(a[k], co[k])); }
    }
    for (int q = 0; q < 3; q++) {
      comz[q] = __shfl_sync(FULL_MASK, z, q); comvz[q] = __shfl_sync(FULL_MASK, vz, q); comaz[q] = __shfl_sync(FULL_MASK, az, q);
    }
    __syncwarp();
  } else {
    for (int q = 0; q < 3; q++) { comz[q] = hcom; comvz[q] = 0; comaz[q] = 0; }
  }
}

// Constraint policy of the warp-cooperative mode: lane r < 24 keeps column r of CI (= -row r of
// A) and ci0_r in registers; the equality column is warp-uniform.
struct StepPolicy {
  double c0, c1, c2, c3, ci0v;   // this lane's constraint (lanes >= 24: unused)
  double e0, e1, e2, e3, ce0v;
  __device__ __forceinline__ double slack(const GiWs& w) const {
    double acc = 0.0;
    acc = fma(c0, w.x[0], acc); acc = fma(c1, w.x[1], acc); acc = fma(c2, w.x[2], acc); acc = fma(c3, w.x[3], acc);
    return acc + ci0v;
  }
  __device__ __forceinline__ void eval_s(const GiWs& w, int lane, double& psi) const {
    if (lane < 24) { const double sv = slack(w); w.s[lane] = sv; psi += fmin(0.0, sv); }
  }
  __device__ __forceinline__ void load_np(const GiWs& w, int ip, int lane, int& klo, int& khi) const {
    const double v0 = __shfl_sync(FULL_MASK, c0, ip), v1 = __shfl_sync(FULL_MASK, c1, ip);
    const double v2 = __shfl_sync(FULL_MASK, c2, ip), v3 = __shfl_sync(FULL_MASK, c3, ip);
    if (lane == 0) { w.np[0] = v0; w.np[1] = v1; w.np[2] = v2; w.np[3] = v3; }
    klo = 0; khi = 4;
    __syncwarp();
  }
  __device__ __forceinline__ double eval_one(const GiWs& w, int ip, int lane) const {
    const double sv = (lane < 24) ? slack(w) : 0.0;
    return __shfl_sync(FULL_MASK, sv, ip);
  }
  __device__ __forceinline__ void load_eq(const GiWs& w, int, int lane, bool& allzero) const {
    if (lane == 0) { w.np[0] = e0; w.np[1] = e1; w.np[2] = e2; w.np[3] = e3; }
    allzero = (fabs(e0) <= 1e-12) && (fabs(e1) <= 1e-12) && (fabs(e2) <= 1e-12) && (fabs(e3) <= 1e-12);
    __syncwarp();
  }
  __device__ __forceinline__ double ce0(int) const { return ce0v; }
};

// Row-pivoted Gauss-Jordan inverse of a 7 x 7 matrix with one lane per column of [A | I]: the same
// element-by-element operation order as gj_inverse7 (bit-identical), 14 lanes wide.
// M: shared memory, row-major 7 x 14, columns 0..6 = A on entry, columns 7..13 = A^-1 on exit.
__device__ void gj_inverse7_warp(double* M, int lane) {
  constexpr int n = 7, W = 14;
  if (lane < W) for (int i = 0; i < n; i++) if (lane >= n) M[i * W + lane] = (lane - n == i) ? 1.0 : 0.0;
  __syncwarp();
  for (int k = 0; k < n; k++) {
    int piv = k;
    double best = fabs(M[k * W + k]);
    for (int i = k + 1; i < n; i++) { const double v = fabs(M[i * W + k]); if (v > best) { best = v; piv = i; } }
    __syncwarp();
    if (piv != k && lane < W) { const double t = M[k * W + lane]; M[k * W + lane] = M[piv * W + lane]; M[piv * W + lane] = t; }
    __syncwarp();
    const double d = M[k * W + k];
    double f[n];
    for (int i = 0; i < n; i++) f[i] = M[i * W + k];
    __syncwarp();
    if (lane < W) {
      const double pk = M[k * W + lane] / d;
      M[k * W + lane] = pk;
      for (int i = 0; i < n; i++) if (i != k) M[i * W + lane] = __dsub_rn(M[i * W + lane], __dmul_rn(f[i], pk));
    }
    __syncwarp();
  }
}

// WARP = false: one THREAD per planner (throughput mode, large batches).
// WARP = true : one WARP per planner (latency mode): the scalar front-end runs warp-uniformly, the
//               QP is solved by the warp-cooperative core (gi_warp.cuh) with one lane per constraint,
//               the 7 x 7 inverse runs one lane per column, lane 0 writes the results.
// 1 block of 128 per SM as the occupancy floor: the compiler then keeps the tick's scalars in 254 registers without
// spills (168 registers by default: 34.5 us per 4096-planner launch with 12 launches in flight, 32.0 us with 254)
#ifndef GO1_STEP_MINB
#define GO1_STEP_MINB 1
#endif
#define GO1_STEP_BOUNDS __launch_bounds__(128, GO1_STEP_MINB)
template <bool WARP>
__global__ void GO1_STEP_BOUNDS step_timing_kernel(StepKParams P) {
  const int lane = threadIdx.x & 31;
  const int b = WARP ? (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5) : (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (b >= P.B) return;
  const bool wr = WARP ? (lane == 0) : true;       // who writes to global memory
  extern __shared__ __align__(16) unsigned char step_smem_raw[];
  double* wsm = reinterpret_cast<double*>(step_smem_raw) + (WARP ? (size_t)(threadIdx.x >> 5) * STEP_WARP_DOUBLES : 0);
  const size_t B = (size_t)P.B;
  const double* S = P.state + b;        // read side
  double* SO = P.state_out + b;         // write side (may alias the read side: in-place update)
  const double* IN = P.in + b;
#define ST(f) S[(size_t)(f) * B]
#define STW(f) if (wr) SO[(size_t)(f) * B]
#define INP(f) IN[(size_t)(f) * B]
  const StepCfgDev& c = P.cfg;
  const double dt = c.dt, Wn = c.Wn;
  const int i = P.tick[b];
  if (i < 1) return;          // no tick for this planner (warp-uniform in the warp-per-planner mapping)
  // The step tables _ts / _tx (27 entries each) are not copied into per-thread arrays (run-time indexed, i.e. local
  // memory, and searched by dependent loads): the searches run over the state's columns as independent coalesced
  // loads, the few entries a tick needs are fetched by index.
  // :702-704 Indexfind((i+1) dt, xyz0 = -1): first entry the time has not passed
  int j = NS;
#pragma unroll
  for (int k = NS - 1; k >= 0; k--) { const double txk = ST(S_TX + k); if (!((i + 1) * dt > txk + 0.0001)) j = k; }
  int p = (j - 1) + 1;
  const bool valid = (p >= 1 && p <= NS);
  if (!valid) p = 1;   // table overrun (UB in the reference): flagged in diag, nothing is written
  const double px = ST(S_FX + p - 1), py = ST(S_FY + p - 1);
  const double tx_p1 = ST(S_TX + p - 1), ts_p1 = ST(S_TS + p - 1), ts_1 = ST(S_TS + 1);
  // what the write-back and CoM_height_solve read of the pre-tick tables, fetched now: two dependent global loads
  // (the period the previous tick left, then its table entries) would otherwise sit exposed behind the SQP loop
  const int bp = (int)ST(S_BJX1);
  const int b1 = bp >= 1 && bp <= NS ? bp : 1, b2 = bp >= 2 && bp <= NS + 1 ? bp : 2;
  const double ts_b1_old = ST(S_TS + b1 - 1), tx_b1_old = ST(S_TX + b1 - 1);
  const double fz_b2 = ST(S_FZ + b2 - 2), fz_b1 = ST(S_FZ + b1 - 1);
  const int ki = (int)round(tx_p1 / dt);
  const int k_yu = i - ki;
  const double Tk = ts_p1 - k_yu * dt;
  const double Lxx_refx = ST(S_LXX + p - 1), Lyy_refy = ST(S_LYY + p - 1);
  const double tr1_ref = cosh_nl(Wn * Tk), tr2_ref = sinh_nl(Wn * Tk);
  double v[4];
  if (i == 1) { v[0] = Lxx_refx; v[1] = Lyy_refy; v[2] = tr1_ref; v[3] = tr2_ref; }
  else { for (int k = 0; k < 4; k++) v[k] = ST(S_VARI + k); }
  double tr1_min, tr2_min;
  if ((c.t_min - k_yu * dt) >= 0.001) { tr1_min = cosh_nl(Wn * (c.t_min - k_yu * dt)); tr2_min = sinh_nl(Wn * (c.t_min - k_yu * dt)); }
  else { tr1_min = cosh_nl(Wn * (0.001)); tr2_min = sinh_nl(Wn * (0.001)); }
  const double tr1_max = cosh_nl(Wn * (c.t_max - k_yu * dt)), tr2_max = sinh_nl(Wn * (c.t_max - k_yu * dt));

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
  double SQ[4][4];
  {
    double SQ0[4][4];
    for (int r = 0; r < 4; r++) for (int k = 0; k < 4; k++) SQ0[r][k] = 0.0;
    SQ0[0][0] = 0.5 * c.bbx;
    SQ0[1][1] = 0.5 * c.bby;
    SQ0[2][2] = 0.5 * (c.rr1 + aax * AxO * AxO + aay * AyO * AyO + aaxv * Axv * Axv + aayv * Ayv * Ayv);
    SQ0[2][3] = 0.5 * (aax * AxO * BxO + aay * AyO * ByO + aaxv * Axv * Bxv + aayv * Ayv * Byv);
    SQ0[3][2] = 0.5 * (aax * BxO * AxO + aay * ByO * AyO + aaxv * Bxv * Axv + aayv * Byv * Ayv);
    SQ0[3][3] = 0.5 * (c.rr2 + aax * BxO * BxO + aay * ByO * ByO + aaxv * Bxv * Bxv + aayv * Byv * Byv);
    for (int r = 0; r < 4; r++) for (int k = 0; k < 4; k++) SQ[r][k] = (SQ0[r][k] + SQ0[k][r]) / 2.0;
  }
  double Sq[4];
  Sq[0] = -c.bbx * Lxx_refx;
  Sq[1] = -c.bby * Lyy_refy;
  Sq[2] = -c.rr1 * tr1_ref + aax * AxO * Cx + aay * AyO * Cy + aaxv * Axv * Cxv + aayv * Ayv * Cyv;
  Sq[3] = -c.rr2 * tr2_ref + aax * BxO * Cx + aay * ByO * Cy + aaxv * Bxv * Cxv + aayv * Byv * Cyv;

  // lateral reachability (:1216-1246)
  double footy_max, footy_min;
  const bool wide = (i >= (round(2 * ts_1 / dt)) + 1);
  const double HW = c.half_hip_width, FW = c.foot_width;
  if (p % 2 == 0) { footy_min = -(2 * HW + 0.03); footy_max = wide ? -(FW + 0.01) : -(HW - 0.03); }
  else { footy_max = 2 * HW + 0.03; footy_min = wide ? FW + 0.01 : HW - 0.03; }

  const double CCx = comx_f - px, CCy = comy_f - py;
  const double sh_dt = c.sh_dt, ch_dt = c.ch_dt;      // sinh / cosh(Wn dt): instance-independent, from the host
  const double AA = Wn * sh_dt;
  const double BBx = (Wn * Wn) * CCx * ch_dt, BBy = (Wn * Wn) * CCy * ch_dt;
  const double AA1x = AA * Wn, AA2x = -2 * AA * CCx * Wn, AA3x = 2 * BBx;
  const double AA1y = AA * Wn, AA2y = -2 * AA * CCy * Wn, AA3y = 2 * BBy;
  const double VAA = ch_dt;
  const double VBBx = Wn * CCx * sh_dt, VBBy = Wn * CCy * sh_dt;
  const double VAA1x = VAA * Wn, VAA2x = -2 * VAA * CCx * Wn, VAA3x = 2 * VBBx - 2 * comvx_f;
  const double VAA1y = VAA * Wn, VAA2y = -2 * VAA * CCy * Wn, VAA3y = 2 * VBBy - 2 * comvy_f;
  const double VAA1x1 = Wn, VAA2x1 = -2 * CCx * Wn, VAA3x1 = -2 * comvx_f;
  const double VAA1y1 = Wn, VAA2y1 = -2 * CCy * Wn, VAA3y1 = -2 * comvy_f;

  int* DG = P.diag ? P.diag + b : nullptr;
#define DGW(f, val) do { if (DG && wr) DG[(size_t)(f) * B] = (val); } while (0)
  int n_solved = 0;
  for (int it = 1; it <= P.n_sqp; it++) {
    double G[16], g0[4];
    for (int r = 0; r < 4; r++) for (int k = 0; k < 4; k++) G[k * 4 + r] = 2 * SQ[r][k];
    for (int r = 0; r < 4; r++) {
      double acc = 0.0;
      for (int k = 0; k < 4; k++) acc += (2 * SQ[r][k]) * v[k];
      g0[r] = acc + Sq[r];
    }
    double CE[4], ce0[1];
    {
      const double trx12[4] = {0.0, 0.0, 2 * v[2], (-2) * v[3]};
      double q = 0.0;
      q += v[2] * v[2];
      q += (v[3] * (-1)) * v[3];
      ce0[0] = -q + 1;
      for (int k = 0; k < 4; k++) CE[k] = trx12[k] * (-1);
    }
    double CI[96], bb[24];   // CI(:, r) = -A(r, :)
#pragma unroll
    for (int k = 0; k < 96; k++) CI[k] = 0.0 * (-1);
#pragma unroll
    for (int k = 0; k < 24; k++) bb[k] = 0.0;
#define AROW(r, k, val) CI[(r) * 4 + (k)] = (val) * (-1)
    AROW(0, 2, 1.0);  bb[0] = -(v[2]) + tr1_max;
    AROW(1, 2, -1.0); bb[1] = -((-1.0) * v[2]) - tr1_min;
    AROW(2, 3, 1.0);  bb[2] = -(v[3]) + tr2_max;
    AROW(3, 3, -1.0); bb[3] = -((-1.0) * v[3]) - tr2_min;
    AROW(4, 0, 1.0);  bb[4] = -(v[0]) + c.footx_max;
    AROW(5, 0, -1.0); bb[5] = -((-1.0) * v[0]) - c.footx_min;
    AROW(6, 1, 1.0);  bb[6] = -(v[1]) + footy_max;
    AROW(7, 1, -1.0); bb[7] = -((-1.0) * v[1]) - footy_min;
    if (k_yu != 0) {
      AROW(8, 0, 1.0);   bb[8] = -(v[0] - Lxx_refx - c.footx_vmax * dt);
      AROW(9, 0, -1.0);  bb[9] = v[0] - Lxx_refx - c.footx_vmin * dt;
      AROW(10, 1, 1.0);  bb[10] = -(v[1] - Lyy_refy - c.footy_vmax * dt);
      AROW(11, 1, -1.0); bb[11] = v[1] - Lyy_refy - c.footy_vmin * dt;
    }
#define ROW3(r, i0, c0, c2, c3, d3) do { \
      AROW(r, i0, (c0)); AROW(r, 2, (c2)); AROW(r, 3, (c3)); \
      double acc_ = 0.0; acc_ += (-(c0)) * v[i0]; acc_ += (-(c2)) * v[2]; acc_ += (-(d3)) * v[3]; bb[r] = acc_; } while (0)
    {
      double c3, d3;
      c3 = AA3x - 2 * c.comax_max; ROW3(12, 0, AA1x, AA2x, c3, c3);
      c3 = -(AA3x - 2 * c.comax_min); ROW3(13, 0, -AA1x, -AA2x, c3, c3);
      c3 = AA3y - 2 * c.comay_max; ROW3(14, 1, AA1y, AA2y, c3, c3);
      c3 = -(AA3y - 2 * c.comay_min); ROW3(15, 1, -AA1y, -AA2y, c3, c3);
      c3 = VAA3x - 2 * c.comax_max * dt; ROW3(16, 0, VAA1x, VAA2x, c3, c3);
      c3 = -(VAA3x - 2 * c.comax_min * dt); ROW3(17, 0, -VAA1x, -VAA2x, c3, c3);
      c3 = VAA3y - 2 * c.comay_max * dt; ROW3(18, 1, VAA1y, VAA2y, c3, c3);
      c3 = -(VAA3y - 2 * c.comay_min * dt); ROW3(19, 1, -VAA1y, -VAA2y, c3, c3);
      c3 = VAA3x1 - 2 * c.comax_max * dt; d3 = VAA3x1 - 2 * c.comax_max * dt / 2.0; ROW3(20, 0, VAA1x1, VAA2x1, c3, d3);
      c3 = -(VAA3x1 - 2 * c.comax_min * dt); d3 = -(VAA3x1 - 2 * c.comax_min * dt / 2.0); ROW3(21, 0, -VAA1x1, -VAA2x1, c3, d3);
      c3 = VAA3y1 - 2 * c.comay_max * dt; d3 = VAA3y1 - 2 * c.comay_max * dt / 2.0; ROW3(22, 1, VAA1y1, VAA2y1, c3, d3);
      c3 = -(VAA3y1 - 2 * c.comay_min * dt); d3 = -(VAA3y1 - 2 * c.comay_min * dt / 2.0); ROW3(23, 1, -VAA1y1, -VAA2y1, c3, d3);
    }
#undef ROW3
#undef AROW
    if (Tk >= 0.1 * ts_p1) {
      double X[4];
      for (int k = 0; k < 4; k++) X[k] = v[k];
      int st, q_iq, q_out, q_add, q_drop, q_degen, qA[5];
      {
        // warp-cooperative solve (same sequence as dense_qp_kernel): lane r < 24 owns constraint r
        GiWs w;
        gi_ws_carve(w, wsm, 4, 1, 24);
        for (int t = lane; t < 4 * w.ld; t += 32) w.J[t] = 0.0;
        if (lane < 16) { const int jj = lane >> 2, ii = lane & 3; w.R[jj * w.ld + ii] = G[lane]; }
        if (lane < 4) w.x[lane] = X[lane];
        const double c1 = ((G[0] + G[5]) + G[10]) + G[15];
        __syncwarp();
        GiResult res; res.f = 0.0; res.iq = 0; res.status = ST_OK;
        res.it_outer = res.it_add = res.it_drop = res.it_degen = res.it_l2a = 0; res.flops = 0;
        if (!gi_llt(w, 4, lane)) {
          res.status = ST_NOT_PD;
        } else {
          gi_inv_lt(w, 4, 0, lane);
          double c2 = (lane < 4) ? w.J[lane * w.ld + lane] : 0.0;
          c2 = warp_sum(c2);
          for (int t = lane; t < 4 * w.ld; t += 32) w.R[t] = 0.0;
          if (lane < 4) w.np[lane] = g0[lane];
          __syncwarp();
          gi_compute_d(w, 0, 4, lane);
          gi_update_z(w, 0, lane);
          double f = 0.0;
          if (lane < 4) { const double xv = -w.z[lane]; w.x[lane] = xv; f = g0[lane] * xv; }
          res.f = 0.5 * warp_sum(f);
          __syncwarp();
          StepPolicy pol;
          const int rr = lane < 24 ? lane : 0;
          pol.c0 = CI[rr * 4 + 0]; pol.c1 = CI[rr * 4 + 1]; pol.c2 = CI[rr * 4 + 2]; pol.c3 = CI[rr * 4 + 3]; pol.ci0v = bb[rr];
          pol.e0 = CE[0]; pol.e1 = CE[1]; pol.e2 = CE[2]; pol.e3 = CE[3]; pol.ce0v = ce0[0];
          gi_loop(w, pol, c1, c2, P.cap, res, lane);
          for (int k = 0; k < 4; k++) X[k] = w.x[k];
          bool has_nan = (X[0] != X[0]) || (X[1] != X[1]) || (X[2] != X[2]) || (X[3] != X[3]);
          if (has_nan && res.status == ST_OK) res.status = ST_NAN;
        }
        st = res.status; q_iq = res.iq; q_out = res.it_outer; q_add = res.it_add; q_drop = res.it_drop; q_degen = res.it_degen;
        for (int k = 0; k < 5; k++) qA[k] = (k < res.iq) ? w.A[k] : 0;
        __syncwarp();
      }
      if (n_solved < STEP_MAX_SQP) {
        const int o = STEP_DIAG_HEAD + n_solved * STEP_DIAG_PER;
        DGW(o + 0, st); DGW(o + 1, st == 1 ? 0 : q_iq);
        DGW(o + 2, q_out); DGW(o + 3, q_add); DGW(o + 4, q_drop); DGW(o + 5, q_degen);
        for (int k = 0; k < 5; k++) DGW(o + 6 + k, (st != 1 && k < q_iq) ? qA[k] : -99);
      }
      n_solved++;
      for (int k = 0; k < 4; k++) v[k] += X[k];   // :795-798, whatever the status
    } else {
      v[0] = Lxx_refx; v[1] = Lyy_refy; v[2] = tr1_ref; v[3] = tr2_ref;
    }
  }
  for (int q = n_solved; q < STEP_MAX_SQP; q++) {      // every diag entry is defined: unused slots read -1, 0...
    DGW(STEP_DIAG_HEAD + q * STEP_DIAG_PER, -1);
    for (int k = 1; k < STEP_DIAG_PER; k++) DGW(STEP_DIAG_HEAD + q * STEP_DIAG_PER + k, 0);
  }

  // write-back (:817, :886-916)
  const double ts_new = k_yu * dt + log(v[2] + v[3]) / Wn;
  const double isx = comx_f - px, esx = v[0] * 0.5, visx = (esx - isx * v[2]) / (1 / Wn * v[3]);
  const double isy = comy_f - py, esy = v[1] * 0.5, visy = (esy - isy * v[2]) / (1 / Wn * v[3]);
  const double fx_next = px + v[0], fy_next = py + v[1];
  if (SO != S) {   // out-of-place: carry the untouched fields over (before any field of the new state is stored)
#pragma unroll 1
    for (int f0 = 0; f0 < STEP_STATE_DOUBLES; f0 += 32) {
      double tmp[32];
#pragma unroll
      for (int k = 0; k < 32; k++) if (f0 + k < STEP_STATE_DOUBLES) tmp[k] = ST(f0 + k);
#pragma unroll
      for (int k = 0; k < 32; k++) if (f0 + k < STEP_STATE_DOUBLES) STW(f0 + k) = tmp[k];
    }
  }
  // _ts(p-1) = ts_new and the running sum _tx(k) = _tx(k-1) + _ts(k-1) for k >= p (:906-909), in the reference's order;
  // in the same pass: the two index searches against the UPDATED table (:1031-1041), the entry CoM_height_solve
  // reads, and the store of the new _tx entries (every load of column k precedes its store: in-place safe)
  const double ts_b1 = (b1 == p) ? ts_new : ts_b1_old;
  double tx_b1 = tx_b1_old;
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
      if (k == b1 - 1) tx_b1 = txk;
      if (jA == NS && !(i * dt >= txk)) jA = k;
      if (jB == NS && !((i + 1) * dt >= txk)) jB = k;
    }
  }

  // vertical CoM samples (:936): after the write-back of ts / tx, with the _bjx1 the previous tick left
  double hz_z[3], hz_vz[3], hz_az[3];
  if (c.ext_height) {
    for (int q = 0; q < 3; q++) { hz_z[q] = INP(I_CZ + q); hz_az[q] = INP(I_CAZ + q); hz_vz[q] = 0.0; }
    hz_vz[0] = INP(I_CVZ);
  } else {
    com_height_solve_warp(i, bp <= NS ? bp : 0, ts_b1, tx_b1, fz_b2, fz_b1, c.hcom, dt, hz_z, hz_vz, hz_az, wsm, lane);
  }
  // LIPM roll-out of samples i, i+1, i+2 (:938-955)
  double comx[3], comy[3], comvx[3], comvy[3], comax[3], comay[3], zmpx[3], zmpy[3], dcmx[3], dcmy[3];
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

  double* O = P.out + b;
#define OUT(f, val) do { if (wr) O[(size_t)(f) * B] = (val); } while (0)
  OUT(0, comx[0]); OUT(1, comy[0]); OUT(2, hz_z[0]);
  OUT(3, comvx[0]); OUT(4, comvy[0]); OUT(5, hz_vz[0]);
  OUT(6, comax[0]); OUT(7, comay[0]); OUT(8, hz_az[0]);
  OUT(9, zmpx[0]); OUT(10, zmpy[0]); OUT(11, dcmx[0]); OUT(12, dcmy[0]);
  OUT(13, zmpx[1]); OUT(14, zmpy[1]); OUT(15, dcmx[1]); OUT(16, dcmy[1]);
  OUT(17, zmpx[2]); OUT(18, zmpy[2]); OUT(19, dcmx[2]); OUT(20, dcmy[2]);
  OUT(21, comax[1]); OUT(22, comay[1]); OUT(23, hz_az[1]);
  OUT(24, comax[2]); OUT(25, comay[2]); OUT(26, hz_az[2]);
  OUT(27, (double)bjxx);
  OUT(28, fx0); OUT(29, fx1); OUT(30, fy0); OUT(31, fy1);
  OUT(32, fz0); OUT(33, fz1);
  OUT(34, (double)(p - 1));
  OUT(35, ts_new);
  OUT(36, v[0]);
  OUT(37, v[1]);
  DGW(0, valid ? p : -1); DGW(1, k_yu); DGW(2, bjxx); DGW(3, bjx1); DGW(4, n_solved);
#undef OUT
#undef DGW
#undef ST
#undef STW
#undef INP
}

cudaError_t step_timing_launch(StepKParams P, cudaStream_t st) {
  const int block = 128, wpc = block / 32;
  const int grid = (P.B + wpc - 1) / wpc;
  step_timing_kernel<true><<<grid, block, (size_t)wpc * STEP_WARP_DOUBLES * sizeof(double), st>>>(P);
  return cudaGetLastError();
}

}  // namespace go1
