// body_split.cu -- body-inclination MPC tick, roll / pitch halves solved side by side.
//
// Same contract as body_fast.cu (replaces PRMPCClass::body_theta_mpc,
// RT/src/FastMPC/PRMPCClass.cpp:379-714, with solve_body_rotation/Solve :799-849 and
// Indexfind :716-738; QP = Eigen::QP::solve_quadprog2, RT/src/utils/EiQuadProg/EiQuadProg.cpp:172-491).
//
// Structure that is exploited.  The QP the reference assembles (cpp:528-565, 799-835) is
//   G  = blockdiag(H, H)                      (H = NH x NH, the same block for roll and pitch),
//   CI = [angle up/low roll | angle up/low pitch | torque up/low roll | torque up/low pitch],
// every constraint column has its support in ONE half.  J = L^-T starts block diagonal, and the
// reference's Givens updates keep every column of J supported in one half (a rotation that meets
// an exact zero of d is a signed swap), so the cross-half entries of d, R and r are exact zeros:
// the reference's solve IS an interleaving of two independent NH-variable solves.  Which half
// moves next is decided only at step 1 / step 2 (EiQuadProg.cpp:282-342): the most negative
// slack over both halves, lowest constraint index on ties.
//
// So: one warp per instance, each 16-lane half-warp runs the Goldfarb-Idnani iteration of its
// half (NH variables, 4 NH constraints) in lock step with the other -- same instruction stream,
// half the matrix sizes, both halves progressing at once -- and logs, per selection, the slack
// value / constraint it picked and, per step-2a pass, what it added or dropped.  A warp-uniform
// replay of the two logs then reproduces the reference's interleaving: the ordered working
// set, the iteration counters and the algorithmic flop count are those of the combined solve.
//
// Anything the replay cannot reproduce from the logs is NOT guessed: a half that ends
// infeasible, degenerate, non-PD, with NaN or over the log capacity, or a final combined
// stopping test that would not have stopped, flags the instance; flagged instances are left
// untouched and appended to a list that a second launch (body_fast_kernel in list mode, the
// combined solve) processes.  On the bench workload the list is empty.
#include <cuda_runtime.h>
#include <stdint.h>
#include "gi_warp.cuh"
#include "tma.cuh"
#include "kernels.h"

#ifndef GO1_SPLIT_WARPS
#define GO1_SPLIT_WARPS 24
#endif

namespace go1 {

namespace {

// The two 16-lane halves of the warp stay CONVERGED: every branch around a warp intrinsic is taken on a
// warp-uniform value and group-specific decisions are predicated, so all shuffles use the compile-time
// full mask (a run-time member mask costs a WARPSYNC per shuffle and serialises REDUX per group).
__device__ __forceinline__ double gbc(double v, int src) { return __shfl_sync(FULL_MASK, v, src, 16); }
__device__ __forceinline__ int gbc(int v, int src) { return __shfl_sync(FULL_MASK, v, src, 16); }

__device__ __forceinline__ double gsum(double v) {
#pragma unroll
  for (int o = 8; o; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ void gsum3(double& a, double& b, double& c) {
#pragma unroll
  for (int o = 8; o; o >>= 1) {
    a += __shfl_xor_sync(FULL_MASK, a, o);
    b += __shfl_xor_sync(FULL_MASK, b, o);
    c += __shfl_xor_sync(FULL_MASK, c, o);
  }
}
// minimum of `v` over the lanes of each half (h = 0 / 1), two full-warp REDUX with the other half neutralised
__device__ __forceinline__ unsigned gmin_u32(int h, unsigned v) {
  const unsigned r0 = __reduce_min_sync(FULL_MASK, h ? 0xffffffffu : v);
  const unsigned r1 = __reduce_min_sync(FULL_MASK, h ? v : 0xffffffffu);
  return h ? r1 : r0;
}
// arg-min over each half: integer reductions on order-preserving keys (see gi_warp.cuh), lowest index on ties
__device__ __forceinline__ void gargmin(int h, double& v, int& idx) {
  unsigned long long k = (unsigned long long)__double_as_longlong(v);
  k ^= (k >> 63) ? 0xffffffffffffffffull : 0x8000000000000000ull;
  const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
  const unsigned mhi = gmin_u32(h, hi);
  const unsigned mlo = gmin_u32(h, hi == mhi ? lo : 0xffffffffu);
  const bool win = (hi == mhi) && (lo == mlo);
  idx = (int)gmin_u32(h, win ? (unsigned)idx : 0xffffffffu);
  unsigned long long m = ((unsigned long long)mhi << 32) | mlo;
  m ^= (m >> 63) ? 0x8000000000000000ull : 0xffffffffffffffffull;
  v = __longlong_as_double((long long)m);
}
// group-masked variants for the rare, group-divergent constraint-drop path
__device__ __forceinline__ double gbcm(unsigned gm, double v, int src) { return __shfl_sync(gm, v, src, 16); }

}  // namespace

template <int NH>
struct SplitDims {
  static constexpr int N = 2 * NH;
  static constexpr int LD = NH | 1;
  static constexpr int JS = (NH * LD + 1) & ~1;
  static constexpr int RP = (NH * (NH + 3) / 2 + 1) & ~1;   // packed R: column c holds rows 0..c+1
  static constexpr int VS = (NH + 1) & ~1;
  static constexpr int OMAX = 16;                           // logged selections per half
  static constexpr int PMAX = 24;                           // logged step-2a passes per half
  static constexpr int GS = JS + RP + 2 * VS + OMAX + OMAX / 2 + PMAX / 2;   // doubles per half
  static constexpr int IN = (36 + 11 * NH + 1) & ~1;
  static constexpr int OUT = (18 + 2 * NH + 1 + 1) & ~1;
  static constexpr int WD = 2 * GS + IN + OUT;              // doubles per warp
  static constexpr int TAB = (3 * NH * NH + 6 * NH + 1) & ~1;
  static_assert(JS >= NH * NH + NH, "the Cholesky factor is staged in J's area");
};

template <int NH, int WPC>
__global__ void __launch_bounds__(WPC * 32, GO1_SPLIT_WARPS / WPC) body_split_kernel(BodyKParams P) {
  using D = SplitDims<NH>;
  constexpr int N = D::N, LD = D::LD, M = 12 * NH;
  static_assert(NH <= 14, "group lanes 14 and 15 carry the scalar divisions of a pass");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* smem = reinterpret_cast<double*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = lane >> 4, gl = lane & 15;
  const unsigned gm = 0xffffu << (h * 16);

  const double* tab = smem;
  const double* ppu = tab;                  // NH x NH, column-major, lower triangular
  const double* gc0 = tab + NH * NH;        // tick-independent part of the Hessian block / 2
  const double* s2 = tab + 2 * NH * NH;     // beta * Ppu'
  const double* m1 = tab + 3 * NH * NH;     // (alpha Pvu') Pvs, NH x 2
  const double* m2 = m1 + 2 * NH;           // (beta Ppu') Pps
  const double* pps = m2 + 2 * NH;          // NH x 2
  double* wbase = smem + D::TAB + (size_t)warp * D::WD;
  double* J = wbase + h * D::GS;            // J(i,j) = J[j*LD + i]   (own half)
  double* Rp = J + D::JS;                   // R(t,c) = Rp[c*(c+3)/2 + t], t <= c+1
  double* xs = Rp + D::RP;                  // broadcast copy of x (g0 during setup)
  double* ds = xs + D::VS;                  // broadcast copy of d
  double* ssv = ds + D::VS;                 // log: slack value of selection o
  int* ipv = reinterpret_cast<int*>(ssv + D::OMAX);          // log: constraint of selection o
  int* plog = ipv + D::OMAX;                                 // log: per pass, added id or 0x10000 | dropped id
  double* inrec = wbase + 2 * D::GS;
  double* outrec = inrec + D::IN;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + D::TAB + (size_t)WPC * D::WD);
  uint64_t* tab_bar = bars + WPC;
  uint64_t* my_bar = bars + warp;

  if (threadIdx.x == 0) {
    for (int i = 0; i <= WPC; i++) mbar_init(bars + i, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(tab_bar, (uint32_t)(D::TAB * sizeof(double)));
    tma_load_1d(smem, P.tab, (uint32_t)(D::TAB * sizeof(double)), tab_bar);
  }
  bool tab_ready = false;
  uint32_t phase = 0;

  const bool act = gl < NH;
  const int kL = act ? gl : 0;
  const double dt = P.dt_mpc, b0 = dt * dt / 2, b1 = dt;
  const double thmax = P.theta_lim, thmin = -P.theta_lim;
  const double j_ini = P.j_ini, tq = P.torque_lim / P.j_ini;
  const double inf = CUDART_INF;

  for (;;) {
    int b = 0;
    if (lane == 0) b = atomicAdd(P.sched, 1);
    b = __shfl_sync(FULL_MASK, b, 0);
    if (b >= P.B) break;

    // ---- stage the input record (one TMA bulk copy) ----
    __syncwarp();
    if (lane == 0) {
      mbar_expect_tx(my_bar, (uint32_t)(D::IN * sizeof(double)));
      tma_load_1d(inrec, P.in + (size_t)b * D::IN, (uint32_t)(D::IN * sizeof(double)), my_bar);
    }
    __syncwarp();
    if (!tab_ready) { mbar_wait(tab_bar, 0); tab_ready = true; }
    mbar_wait(my_bar, phase);
    phase ^= 1u;

    double* outg = P.out + (size_t)b * D::OUT;
    const int tick = (int)inrec[27];
    const double thx0 = inrec[28], thx1 = inrec[29], thy0 = inrec[30], thy1 = inrec[31];
    const double xw = act ? inrec[36 + h * NH + kL] : 0.0;   // warm start entry (kept on a gated tick)
    const double* refs = inrec + 36 + N;

    int bjx1 = 0, bjx2 = 0, status = -1;
    int iqc = 0, it_outer = 0, it_add = 0, it_drop = 0, it_l2a = 0;
    unsigned flops = 0;
    int Ac = 0;              // slot `lane` of the combined working set
    bool flag = false;       // hand the instance to the combined kernel
    double f_value = 0.0;

    bool live = false;
    int i = tick;
    if (!(i < P.gate)) { i -= P.gate; live = (i < P.nsum_mpc - NH); }

    if (!live) {
      // gated tick: the reference returns its stale members; state and V_ini unchanged
      double o = (lane < 14) ? outg[lane] : 0.0;
      double thl = (lane < 4) ? inrec[28 + lane] : 0.0;
      __syncwarp();
      if (lane < 14) outrec[lane] = o;
      if (lane < 4) outrec[14 + lane] = thl;
      if (act) outrec[18 + h * NH + kL] = xw;
      if (lane == 0) outrec[18 + N] = 0.0;
    } else {
      // ---- phase indices (cpp:406-417): first table entry the time has not reached ----
      {
        double txl = (lane < 27) ? inrec[lane] : inf;
        unsigned g1 = __ballot_sync(FULL_MASK, !((i + 1) * dt >= txl));
        unsigned g2 = __ballot_sync(FULL_MASK, !((i + NH) * dt >= txl));
        bjx1 = __ffs(g1) - 1;
        bjx2 = __ffs(g2) - 1;
      }
      const int t_yu = (i + 1) % P.nstepx;
      const bool left = (bjx1 < 2) || (bjx1 % 2 == 0);
      const bool sw = (bjx1 >= 2) && !((t_yu + NH - 1) < P.nstepx);
      const int t_yu_k = (t_yu + NH) - P.nstepx;

      // ---- condensation (cpp:427-526): group lane k owns horizon step k of half h ----
      const double my0 = h ? thy0 : thx0, my1 = h ? thy1 : thx1;
      double g0 = 0.0, pk = 0.0, pth = 0.0;
      double zx_o = 0.0, zy_o = 0.0, caz_o = 0.0;            // lanes 0..2: outputs need steps 0..2
      const double bs_p = (lane < 2) ? inrec[32 + 2 * lane] : 0.0, bs_v = (lane < 2) ? inrec[33 + 2 * lane] : 0.0;
      if (lane < 3) { zx_o = refs[lane]; zy_o = refs[NH + lane]; caz_o = refs[8 * NH + lane]; }
      if (act) {
        const int k = kL;
        const bool other = sw && (k >= NH - t_yu_k);
        const bool use_l = left ? !other : other;
        const double cop = h ? (use_l ? refs[6 * NH + k] : refs[4 * NH + k])      // half 1 (pitch accel) uses det_px
                             : (use_l ? refs[7 * NH + k] : refs[5 * NH + k]);     // half 0 (roll accel) uses det_py
        const double det = (h ? refs[k] : refs[NH + k]) - cop;
        pth = j_ini / (P.mass * (refs[8 * NH + k] + P.g));
        pk = fma(pps[k], my0, pps[NH + k] * my1);
        const double t1 = fma(m1[k], my0, m1[NH + k] * my1), t2 = fma(m2[k], my0, m2[NH + k] * my1);
        const double* bref = refs + (2 + h) * NH;
        double t3 = 0.0;
#pragma unroll
        for (int j = 0; j < NH; j++) t3 = fma(s2[j * NH + k], bref[j], t3);
        g0 = ((t1 + t2) - t3) + (P.gama * (h ? -pth : pth)) * det;
      }

      // ---- Hessian block, right-looking Cholesky: group lanes own rows, the finished column goes through
      //      shared memory (both halves factor the same block, each into its own J area) ----
      double* Ls = J;             // L(i,k) = Ls[k*NH + i]
      double* linvs = J + NH * NH;
      double Lrow[NH];
      double tr = 0.0;
#pragma unroll
      for (int j = 0; j < NH; j++) {
        double v = act ? gc0[j * NH + gl] : 0.0;
        if (j == gl) { v = v + P.gama / 2 * (pth * pth); tr = 2 * v; }
        Lrow[j] = 2 * v;
      }
      tr = gsum(tr);
      const double c1 = 2 * tr;
      bool pd = true;
      double c2 = 0.0;
#pragma unroll
      for (int k = 0; k < NH; k++) {
        const double piv = gbc(Lrow[k], k);
        if (!(piv > 0.0)) { pd = false; break; }   // <= 0 or NaN: the combined kernel deals with it (warp-uniform)
        const double rs = rsqrt(piv);
        const double lk = (gl >= k) ? Lrow[k] * rs : 0.0;
        c2 += rs;
        if (act) Ls[k * NH + gl] = lk;
        if (gl == k) linvs[k] = rs;
        __syncwarp();
#pragma unroll
        for (int j = k + 1; j < NH; j++) Lrow[j] = fma(-lk, Ls[k * NH + j], Lrow[j]);
      }

      double x = xw, u = 0.0, rinv = 0.0;
      int A = 0, iq = 0, nout = 0, npass = 0;
      double psi_end = 0.0;
      bool end_tol = false;          // the half stopped on |psi| <= tol (it may still hold candidates)
      double R_norm = 1.0, dq_min = inf;
      double tol = 0.0;
      if (!pd) {
        flag = true;
      } else {
        // J = L^-T: every group lane builds column kL by back substitution
        if (act) xs[gl] = g0;
        __syncwarp();
        double y[NH];
#pragma unroll
        for (int ii = NH - 1; ii >= 0; ii--) {
          double t = 0.0;
#pragma unroll
          for (int k = ii + 1; k < NH; k++) t = fma(Ls[ii * NH + k], y[k], t);
          const double li = linvs[ii];
          y[ii] = (ii == kL) ? li : ((ii < kL) ? -t * li : 0.0);
        }
        c2 = 2 * c2;
        // d0 = J' g0 (column owner has the column in registers), columns to shared memory
        double d0 = 0.0;
#pragma unroll
        for (int ii = 0; ii < NH; ii++) d0 = fma(y[ii], xs[ii], d0);
        __syncwarp();
        if (act) {
          double* col = J + gl * LD;
#pragma unroll
          for (int ii = 0; ii < NH; ii++) col[ii] = y[ii];
          ds[gl] = d0;
        }
        __syncwarp();
        // x = -J d0, f = g0.x / 2
        {
          double acc = 0.0;
          if (act) {
            double a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int j = 0; j < NH; j += 2) {
              a0 = fma(J[j * LD + gl], ds[j], a0);
              if (j + 1 < NH) a1 = fma(J[(j + 1) * LD + gl], ds[j + 1], a1);
            }
            acc = a0 + a1;
          }
          x = -acc;
          f_value = 0.5 * gsum(act ? g0 * x : 0.0);
        }
        __syncwarp();

        // ================= Goldfarb-Idnani iteration of this half (EiQuadProg.cpp:282-490) =================
        tol = M * EPS_D * c1 * c2 * 100.0;
        unsigned inA = 0u;                            // bit s <-> this lane's constraint slot s
                // constraint index of slot s of this lane: blk = ((s>>1)<<2) | (h<<1) | (s&1), id = blk NH + k
        const int cbase = (h << 1) * NH + kL;
        bool fin = false, need_l1 = true;
        int ip = 0, ip_blk = 0, ip_k = 0;
        double sip = 0.0, npL = 0.0, ip_sgn = 1.0;

        auto slack_angle = [&](double& up, double& low) {
          double v0 = 0.0, v1 = 0.0;
#pragma unroll
          for (int j = 0; j < NH; j += 2) {
            if (j <= kL) v0 = fma(ppu[j * NH + kL], xs[j], v0);
            if (j + 1 < NH && j + 1 <= kL) v1 = fma(ppu[(j + 1) * NH + kL], xs[j + 1], v1);
          }
          const double v = v0 + v1;
          up = (thmax - pk) - v;
          low = v + (thmax + pk);
        };

        for (;;) {
          if (__all_sync(FULL_MASK, fin)) break;
          const bool g1 = need_l1 && !fin;
          if (__any_sync(FULL_MASK, g1)) {
            // ---- step 1 (cpp:282-320) and step 2 (cpp:322-342) for the halves that need them ----
            if (act && g1) xs[gl] = x;
            __syncwarp();
            double n0 = 0, n1 = 0, n2 = 0, n3 = 0, psi = 0.0;
            if (act) {
              slack_angle(n0, n1);
              n2 = fma(-j_ini, x, tq);
              n3 = fma(j_ini, x, tq);
              psi = (fmin(0.0, n0) + fmin(0.0, n1)) + (fmin(0.0, n2) + fmin(0.0, n3));
            }
            psi = gsum(psi);
            // most negative eligible slack, lowest constraint index among equals
            double bv = 0.0; int bi = 0x7fffffff;
            if (act) {
              if (!(inA & 1u) && n0 < bv) { bv = n0; bi = cbase; }
              if (!(inA & 2u) && n1 < bv) { bv = n1; bi = cbase + NH; }
              if (!(inA & 4u) && n2 < bv) { bv = n2; bi = cbase + 4 * NH; }
              if (!(inA & 8u) && n3 < bv) { bv = n3; bi = cbase + 5 * NH; }
            }
            gargmin(h, bv, bi);
            if (g1) {
              psi_end = psi;
              if (!(fabs(psi) > tol)) { fin = true; end_tol = true; if (psi != psi) flag = true; }
              else if (!(bv < 0.0)) fin = true;
              else if (nout >= D::OMAX) { fin = true; flag = true; }
              else {
                ip = bi; sip = bv;
                if (gl == 0) { ssv[nout] = bv; ipv[nout] = bi; }
                nout++;
                ip_blk = ip / NH; ip_k = ip - ip_blk * NH;
                ip_sgn = (ip_blk & 1) ? 1.0 : -1.0;
                npL = 0.0;
                if (act) {
                  if (ip_blk < 4) { if (kL <= ip_k) npL = ip_sgn * ppu[kL * NH + ip_k]; }
                  else if (kL == ip_k) npL = ip_sgn * j_ini;
                }
                if (gl == iq) { u = 0.0; A = ip; }
                need_l1 = false;
              }
            }
          }
          if (!fin && npass >= D::PMAX) { fin = true; flag = true; }
          const bool run = !fin;
          if (!__any_sync(FULL_MASK, run)) continue;
          // ---- step 2a (cpp:349-386); a finished half computes along and commits nothing ----
          // d = J' n+ : group lane owns column gl; n+ comes from the model table
          double d = 0.0;
          if (act) {
            const double* col = J + gl * LD;
            if (ip_blk < 4) {
              double a0 = 0.0, a1 = 0.0;
              int ii = 0;
              if (!(ip_k & 1)) { a0 = col[0] * ppu[ip_k]; ii = 1; }     // ip_k + 1 terms: peel one when odd
#pragma unroll 2
              for (; ii <= ip_k; ii += 2) {
                a0 = fma(col[ii], ppu[ii * NH + ip_k], a0);
                a1 = fma(col[ii + 1], ppu[(ii + 1) * NH + ip_k], a1);
              }
              d = ip_sgn * (a0 + a1);
            } else {
              d = ip_sgn * (j_ini * col[ip_k]);
            }
            ds[gl] = d;
          }
          __syncwarp();
          // z = J[:, iq:] d[iq:]
          double z = 0.0;
          if (act) {
            double z0 = 0.0, z1 = 0.0;
            int j = iq;
            if ((NH - j) & 1) { z0 = J[j * LD + gl] * ds[j]; j++; }
#pragma unroll 2
            for (; j < NH; j += 2) {
              z0 = fma(J[j * LD + gl], ds[j], z0);
              z1 = fma(J[(j + 1) * LD + gl], ds[j + 1], z1);
            }
            z = z0 + z1;
          }
          // r = R^-1 d[0:iq)  (column-oriented back substitution, stored reciprocals); the trip count is the
          // larger working set of the two halves, the shorter one idles through predication
          double r = (gl < iq) ? d : 0.0;
          {
            const int iqo = __shfl_xor_sync(FULL_MASK, iq, 16);
            for (int c = max(iq, iqo) - 1; c >= 0; c--) {
              const double rc = gbc((gl < iq) ? r * rinv : 0.0, c);
              if (c < iq) {
                if (gl == c) r = rc;
                else if (gl < c) r = fma(-rc, Rp[c * (c + 3) / 2 + gl], r);
              }
            }
          }
          // step lengths: ONE division instruction serves the ratio test (lanes < iq: u/r), t2 (group
          // lane 15: -s_ip / z.n+) and the Householder scale of a possible add (group lane 14)
          double zz = z * z, zn = z * npL, dd = (act && gl >= iq) ? d * d : 0.0;
          gsum3(zz, zn, dd);
          const double inrm = (dd > 0.0) ? rsqrt(dd) : 0.0;      // 1 / |d2|
          const double nrm = dd * inrm;                           // |d2|
          const double diq = gbc(d, iq);
          double num = u, den = r;
          if (gl == 15) { num = -sip; den = zn; }
          if (gl == 14) { num = 1.0; den = nrm * (nrm + fabs(diq)); }
          const double quo = num / den;
          double t1 = inf; int kmin = 0x7fffffff;
          if (gl < iq && r > 0.0) { t1 = quo; kmin = gl; }
          gargmin(h, t1, kmin);
          const int Al = gbc(A, kmin & 15);
          const int l = (kmin != 0x7fffffff && t1 < inf) ? Al : 0;
          const double q15 = gbc(quo, 15);
          const double t2 = (fabs(zz) > EPS_D) ? q15 : inf;
          const double tau = gbc(quo, 14);
          const double uiq = gbc(u, iq);
          const double t = fmin(t1, t2);
          const bool go = run && (t < inf);
          const bool prim = go && !(t2 >= inf);                  // case (iii); go && !prim: case (ii), dual step
          const bool full = prim && (t == t2);
          const bool dropg = go && !full;
          if (run && !go) { fin = true; flag = true; }           // case (i): infeasible (or NaN)
          if (go) {
            if (gl < iq) u = fma(-t, r, u);
            if (gl == iq) u += t;
          }
          if (prim) {
            x = fma(t, z, x);
            f_value += t * zn * (0.5 * t + uiq);
          }
          if (full) {
            // ---- add_constraint (cpp:30-93) as ONE Householder reflection (see body_fast.cu) ----
            const double sigma = (diq < 0.0) ? -nrm : nrm;
            if (nrm != 0.0 && act) {
              const double sw2 = tau * fma(sigma, J[iq * LD + gl], z);   // tau * (J2 v)_k
              const double viq = diq + sigma;
              J[iq * LD + gl] = fma(-sw2, viq, J[iq * LD + gl]);
#pragma unroll 4
              for (int j = iq + 1; j < NH; j++) J[j * LD + gl] = fma(-sw2, ds[j], J[j * LD + gl]);
            }
            if (gl == iq) d = (nrm != 0.0) ? -sigma : d;
            if (gl <= iq) Rp[iq * (iq + 3) / 2 + gl] = d;
            if (gl == iq) rinv = (nrm != 0.0) ? ((diq < 0.0) ? inrm : -inrm) : 1.0 / d;   // 1 / (-sigma)
            const double dq = (nrm != 0.0) ? -sigma : diq;           // new R(iq,iq), group-uniform
            iq++;
            if (gl == 0) plog[npass] = ip;
            npass++;
            if (fabs(dq) <= EPS_D * R_norm) { fin = true; flag = true; }   // degenerate: combined kernel
            R_norm = fmax(R_norm, fabs(dq));
            dq_min = fmin(dq_min, fabs(dq));
            {
              const int slot = ((ip_blk >> 2) << 1) | (ip_blk & 1);
              if (gl == ip_k) inA |= 1u << slot;
            }
            need_l1 = true;
          }
          __syncwarp();
          if (!__any_sync(FULL_MASK, dropg)) continue;
          // ---- delete_constraint(l) (cpp:95-170) after a dual step or a partial step: rare (0.2 per solve),
          //      the one group-divergent region; member-masked intrinsics, reconverged behind it ----
          if (dropg) {
            const bool dual = !prim;
            {
              const int lblk = l / NH, lk = l - lblk * NH;
              const int slot = ((lblk >> 2) << 1) | (lblk & 1);
              if (gl == lk) inA &= ~(1u << slot);
            }
            const unsigned hit = __ballot_sync(gm, gl < iq && A == l) >> (h * 16);
            if (!hit) { fin = true; flag = true; }   // l not in the working set: UB in the reference
            else {
              const int qq = __ffs(hit) - 1;
              {
                const int An = __shfl_down_sync(gm, A, 1, 16);
                const double un = __shfl_down_sync(gm, u, 1, 16);
                if (gl >= qq && gl < iq) { A = An; u = un; }
                if (gl == iq) { A = 0; u = 0.0; }
              }
              // R columns shift left by one (column c+1 -> c, rows 0..c+1)
              for (int c = qq; c < iq - 1; c++) {
                const double v = (gl <= c + 1) ? Rp[(c + 1) * (c + 4) / 2 + gl] : 0.0;
                __syncwarp(gm);
                if (gl <= c + 1) Rp[c * (c + 3) / 2 + gl] = v;
              }
              iq--;
              __syncwarp(gm);
              if (gl == 0) plog[npass] = 0x10000 | l;
              npass++;
              for (int j = qq; j < iq; j++) {
                double cc = Rp[j * (j + 3) / 2 + j], sn = Rp[j * (j + 3) / 2 + j + 1];
                const double hh = gi_hypot(cc, sn);
                __syncwarp(gm);
                if (hh == 0.0) continue;
                cc = cc / hh; sn = sn / hh;
                const double diag = (cc < 0.0) ? -hh : hh;
                if (gl == 0) { Rp[j * (j + 3) / 2 + j + 1] = 0.0; Rp[j * (j + 3) / 2 + j] = diag; }
                if (cc < 0.0) { cc = -cc; sn = -sn; }
                const double xny = sn / (1.0 + cc);
                if (gl > j && gl < iq) {
                  double* cp = Rp + gl * (gl + 3) / 2;
                  const double t1j = cp[j], t2j = cp[j + 1];
                  const double a = fma(t2j, sn, t1j * cc);
                  cp[j] = a;
                  cp[j + 1] = fma(xny, t1j + a, -t2j);
                }
                if (act) {
                  const double t1j = J[j * LD + gl], t2j = J[(j + 1) * LD + gl];
                  const double a = fma(t2j, sn, t1j * cc);
                  J[j * LD + gl] = a;
                  J[(j + 1) * LD + gl] = fma(xny, a + t1j, -t2j);
                }
                __syncwarp(gm);
              }
              if (gl >= qq && gl < iq) rinv = 1.0 / Rp[gl * (gl + 3) / 2 + gl];
              if (!dual) {
                // partial step: recompute the slack of ip at the new x (a dual step keeps s_ip)
                if (act) xs[gl] = x;
                __syncwarp(gm);
                double sv = 0.0;
                if (gl == ip_k) {
                  if (ip_blk < 4) { double up, low; slack_angle(up, low); sv = (ip_blk & 1) ? low : up; }
                  else sv = (ip_blk & 1) ? fma(j_ini, x, tq) : fma(-j_ini, x, tq);
                }
                sip = gbcm(gm, sv, ip_k);
              }
            }
          }
          __syncwarp();
        }
      }

      // ---- replay of the combined iteration from the two logs (warp-uniform) ----
      __syncwarp();
      flag = __any_sync(FULL_MASK, flag || (act && (x != x)));
      if (!flag) {
        const int nout_x = __shfl_sync(FULL_MASK, nout, 0), nout_y = __shfl_sync(FULL_MASK, nout, 16);
        const double psi_x = __shfl_sync(FULL_MASK, psi_end, 0), psi_y = __shfl_sync(FULL_MASK, psi_end, 16);
        const bool tol_x = __shfl_sync(FULL_MASK, (int)end_tol, 0), tol_y = __shfl_sync(FULL_MASK, (int)end_tol, 16);
        const double rn = fmax(__shfl_sync(FULL_MASK, R_norm, 0), __shfl_sync(FULL_MASK, R_norm, 16));
        const double dqm = fmin(__shfl_sync(FULL_MASK, dq_min, 0), __shfl_sync(FULL_MASK, dq_min, 16));
        // the combined solve stops when |psi_x + psi_y| <= tol or nothing is eligible; a half that stopped on its
        // own tolerance may still hold a (tiny) candidate the combined scan would pick: do not guess.
        if (fabs(psi_x + psi_y) > tol && ((tol_x && psi_x != 0.0) || (tol_y && psi_y != 0.0))) flag = true;
        if (dqm <= EPS_D * rn) flag = true;   // an add the combined degeneracy test (max R_norm over both halves) could reject
        const double* ssx = wbase + D::JS + D::RP + 2 * D::VS;
        const double* ssy = ssx + D::GS;
        const int* ipx = reinterpret_cast<const int*>(ssx + D::OMAX);
        const int* ipy = reinterpret_cast<const int*>(ssy + D::OMAX);
        const int* plx = ipx + D::OMAX;
        const int* ply = ipy + D::OMAX;
        int ox = 0, oy = 0, px = 0, py = 0;
        flops = (unsigned)gi_flops_setup(N, 0);
        for (;;) {
          it_outer++;
          flops += 2u * N * M;
          const bool hx = ox < nout_x, hy = oy < nout_y;
          if (!hx && !hy) break;
          bool pickx = hx;
          if (hx && hy) {
            const double a = ssx[ox], c = ssy[oy];
            pickx = (a < c) || (a == c && ipx[ox] < ipy[oy]);
          }
          for (;;) {
            const int e = pickx ? plx[px++] : ply[py++];
            it_l2a++;
            flops += 2u * N * N + 2u * N * (N - iqc) + (unsigned)(iqc * iqc) + 4u * N + 2u * iqc;
            if (!(e & 0x10000)) {
              flops += 6u * N * (unsigned)(N - iqc - 1 > 0 ? N - iqc - 1 : 0);
              if (lane == iqc) Ac = e;
              iqc++; it_add++;
              break;
            }
            const int l = e & 0xffff;
            const unsigned hit = __ballot_sync(FULL_MASK, lane < iqc && Ac == l);
            const int qq = __ffs(hit) - 1;
            const int An = __shfl_down_sync(FULL_MASK, Ac, 1);
            if (lane >= qq && lane < iqc) Ac = An;
            iqc--; it_drop++;
            flops += 3u * (unsigned)((iqc - qq) * (iqc - qq)) + 6u * N * (unsigned)(iqc - qq);
          }
          if (pickx) ox++; else oy++;
        }
        status = ST_OK;
        f_value = __shfl_sync(FULL_MASK, f_value, 0) + __shfl_sync(FULL_MASK, f_value, 16);
      }

      if (!flag) {
        // ---- first control: clamp (cpp:567-625); NaN never reaches this point ----
        double ax0 = __shfl_sync(FULL_MASK, x, 0), ay0 = __shfl_sync(FULL_MASK, x, 16);
        const double arow_x = thx0 + dt * thx1, arow_y = thy0 + dt * thy1;
        {
          const double nx0 = arow_x + b0 * ax0;
          if (nx0 > thmax) ax0 = (thmax - arow_x) / b0;
          else if (nx0 < thmin) ax0 = (thmin - arow_x) / b0;
          const double ny0 = arow_y + b0 * ay0;
          if (ny0 > thmax) ay0 = (thmax - arow_y) / b0;
          else if (ny0 < thmin) ay0 = (thmin - arow_y) / b0;
        }
        if (lane == 0) x = ax0;
        if (lane == 16) x = ay0;
        // ---- roll-out (cpp:629-655): lane 0 integrates roll, lane 1 pitch; out14 needs steps 0..2 ----
        const double xa1 = __shfl_sync(FULL_MASK, x, 1), xa2 = __shfl_sync(FULL_MASK, x, 2);
        const double ya1 = __shfl_sync(FULL_MASK, x, 17), ya2 = __shfl_sync(FULL_MASK, x, 18);
        if (lane < 2) {
          const double a0 = lane ? ay0 : ax0, a1 = lane ? ya1 : xa1, a2 = lane ? ya2 : xa2;
          const double p0 = lane ? thy0 : thx0, v0 = lane ? thy1 : thx1;
          const double lam_p = P.lamda[2 * lane], lam_v = P.lamda[2 * lane + 1];
          double pkk = (p0 + dt * v0) + b0 * a0, vk = v0 + b1 * a0;
          outrec[14 + 2 * lane] = lam_p * bs_p + (1 - lam_p) * pkk;
          outrec[15 + 2 * lane] = lam_v * bs_v + (1 - lam_v) * vk;
          outrec[0 + lane] = pkk;
          double pn = (pkk + dt * vk) + b0 * a1; vk = vk + b1 * a1; pkk = pn;
          outrec[6 + lane] = pkk;
          pn = (pkk + dt * vk) + b0 * a2; pkk = pn;
          outrec[10 + lane] = pkk;
          outrec[2 + lane] = j_ini * a0;
        }
        {
          // cpp:651-652 ZMP consistent with the planned angular acceleration (steps 0..2)
          const double xacc = x;                                        // lane jj < 3: roll accel at step jj
          const double yacc = __shfl_sync(FULL_MASK, x, (16 + lane) & 31);   // pitch accel at step jj
          if (lane < 3) {
            const double den = P.mass * (P.g + caz_o);
            const int o = (lane == 0) ? 4 : (lane == 1 ? 8 : 12);
            outrec[o] = zx_o - j_ini * yacc / den;
            outrec[o + 1] = zy_o + j_ini * xacc / den;
          }
        }
        if (act) outrec[18 + h * NH + kL] = x;
        if (lane == 0) outrec[18 + N] = f_value;
      }
    }
    if (flag) {
      // leave the record untouched; the combined kernel (list mode) takes the instance
      if (lane == 0) {
        const int slot = atomicAdd(P.flist_count, 1);
        if (slot < P.flist_cap) P.flist[slot] = b;
      }
      __syncwarp();
      continue;
    }
    if (lane == 0 && D::OUT > 19 + N) outrec[19 + N] = 0.0;   // pad double of the record: defined, not stale shared memory
    // ---- write back: one TMA bulk store of the output record, diagnostics by lanes ----
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      tma_store_1d(outg, outrec, (uint32_t)(D::OUT * sizeof(double)));
      tma_store_commit();
    }
    if (P.diag) {
      int* dg = P.diag + (size_t)b * P.diag_stride;
      if (lane == 0) {
        dg[0] = status; dg[1] = iqc;
        dg[2] = it_outer; dg[3] = it_add; dg[4] = it_drop; dg[5] = 0;
        dg[6] = bjx1; dg[7] = bjx2; dg[8] = it_l2a; dg[9] = (int)flops;
      }
      if (lane < N) dg[10 + lane] = (lane < iqc) ? Ac : -1;
    }
    if (lane == 0) tma_store_wait_read();
    __syncwarp();
  }
  if (lane == 0) {
    tma_store_wait_all();
    // last warp out resets the instance counter for the next launch
    const int done = atomicAdd(P.sched + 1, 1);
    if (done == (int)(gridDim.x * WPC) - 1) { P.sched[0] = 0; P.sched[1] = 0; __threadfence(); }
  }
}

template <int NH>
static cudaError_t split_launch_nh(const BodyKParams& P, int sms, cudaStream_t st) {
  constexpr int WPC = 4;
  using D = SplitDims<NH>;
  if (P.in_stride != D::IN || P.out_stride != D::OUT || P.tab_doubles != D::TAB) return cudaErrorInvalidValue;
  const size_t smem = (size_t)(D::TAB + WPC * D::WD) * sizeof(double) + (size_t)(WPC + 1) * sizeof(uint64_t);
  int dev_ = 0;            // function attributes and occupancy are per device
  cudaGetDevice(&dev_);
  static int occ_cache_[64] = {};
  int& occ_cache = occ_cache_[dev_ & 63];
  if (occ_cache == 0) {
    cudaError_t e = cudaFuncSetAttribute(body_split_kernel<NH, WPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_cache, body_split_kernel<NH, WPC>, WPC * 32, smem);
    if (e != cudaSuccess) return e;
    if (occ_cache < 1) return cudaErrorLaunchOutOfResources;
  }
  int grid = (P.B + WPC - 1) / WPC;
  if (grid > sms * occ_cache) grid = sms * occ_cache;
  body_split_kernel<NH, WPC><<<grid, WPC * 32, smem, st>>>(P);
  return cudaGetLastError();
}

bool body_split_supported(int nh) { return nh == 10; }

cudaError_t body_split_launch(BodyKParams P, int sms, cudaStream_t st) {
  switch (P.nh) {
    case 10: return split_launch_nh<10>(P, sms, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace go1
