// body_fast.cu -- body-inclination MPC tick, compile-time horizon, register-resident solver state.
//
// Same contract as body_mpc.cu (replaces PRMPCClass::body_theta_mpc,
// RT/src/FastMPC/PRMPCClass.cpp:379-714, with solve_body_rotation/Solve :799-849 and
// Indexfind :716-738; QP = Eigen::QP::solve_quadprog2, RT/src/utils/EiQuadProg/EiQuadProg.cpp:172-491)
// for horizons NH <= 15, i.e. n = 2 NH <= 30 variables: one warp per instance, and
//   * lane L < n owns variable L: x, z, n+ entry, gradient entry, the four constraints that
//     bound it (angle up/low at its horizon step, torque up/low) with their slacks and
//     working-set bits -- all in registers;
//   * lane t owns slot t of the working set (A, u, r, 1/R_tt) in registers; deleting a
//     constraint is a shuffle-down;
//   * lane j owns column j of J for d = J' n+ and row L of J for z = J2 d2 and the Givens
//     sweeps; J (n x n, odd leading dimension) and the packed upper-triangular R are the
//     only solver state in shared memory (6.5 KB per warp at NH = 10 -> 28 warps/SM);
//   * n+ and the slacks come straight from the CTA-shared horizon model (Ppu), the QP
//     matrices G, CI, ci0 never exist;
//   * every loop bound is a compile-time constant (NH), triangular solves multiply by stored
//     reciprocals, the Cholesky pivot uses rsqrt: no division or square root on the
//     per-iteration critical path except the step-length ratio test and the Givens setup;
//   * instances are handed out by an atomic counter (iteration counts vary 1..n per
//     instance), which self-resets at the end of the launch.
// Rounding differs from the CPU oracle in the last bits; parity (1e-9 relative on x,
// identical active set and iteration counters) is checked by tests/test_gpu_body.py.
#include <cuda_runtime.h>
#include <mutex>
#include <stdint.h>
#include "gi_warp.cuh"
#include "tma.cuh"
#include "kernels.h"

#ifndef GO1_FAST_WARPS
#define GO1_FAST_WARPS 16   // resident warps per SM the NH <= 10 instantiations are register-limited to (128 regs: no spills; measured faster than 20/24/28)
#endif

namespace go1 {

__device__ __forceinline__ double bcast(double v, int src) { return __shfl_sync(FULL_MASK, v, src); }

// two warp sums at once (independent butterflies interleave)
__device__ __forceinline__ void warp_sum2(double& a, double& b) {
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    a += __shfl_xor_sync(FULL_MASK, a, o);
    b += __shfl_xor_sync(FULL_MASK, b, o);
  }
}

__device__ __forceinline__ void warp_sum3(double& a, double& b, double& c) {
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    a += __shfl_xor_sync(FULL_MASK, a, o);
    b += __shfl_xor_sync(FULL_MASK, b, o);
    c += __shfl_xor_sync(FULL_MASK, c, o);
  }
}

template <int NH>
struct FastDims {
  static constexpr int N = 2 * NH;
  static constexpr int LD = N | 1;
  static constexpr int RP = (N * (N + 3) / 2 + 1) & ~1;   // packed R: column c holds rows 0..c+1
  static constexpr int IN = (36 + 11 * NH + 1) & ~1;
  static constexpr int OUT = (18 + 2 * NH + 1 + 1) & ~1;
  static constexpr int JS = (N * LD + 1) & ~1;
  static constexpr int WD = JS + RP + IN + OUT;           // doubles per warp
  static constexpr int TAB = (3 * NH * NH + 6 * NH + 1) & ~1;
};

template <int NH, int WPC>
__global__ void __launch_bounds__(WPC * 32, (NH <= 10 ? GO1_FAST_WARPS : 12) / WPC) body_fast_kernel(BodyKParams P) {
  using D = FastDims<NH>;
  constexpr int N = D::N, LD = D::LD, M = 12 * NH;
  static_assert(N <= 28, "lanes 30 and 31 carry the scalar divisions of a pass: the working set must stay below them");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* smem = reinterpret_cast<double*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const double* tab = smem;
  const double* ppu = tab;                  // NH x NH, column-major, lower triangular
  const double* gc0 = tab + NH * NH;        // tick-independent part of the Hessian block / 2
  const double* s2 = tab + 2 * NH * NH;     // beta * Ppu'
  const double* m1 = tab + 3 * NH * NH;     // (alpha Pvu') Pvs, NH x 2
  const double* m2 = m1 + 2 * NH;           // (beta Ppu') Pps
  const double* pps = m2 + 2 * NH;          // NH x 2
  double* wbase = smem + D::TAB + (size_t)warp * D::WD;
  double* J = wbase;                        // J(i,j) = J[j*LD + i]
  double* Rp = J + D::JS;                   // R(t,c) = Rp[c*(c+3)/2 + t], t <= c+1
  double* inrec = Rp + D::RP;
  double* outrec = inrec + D::IN;
  // scratch overlays the input record once its fields sit in registers
  double* xs = inrec;                       // N: broadcast copy of x (also g0 during setup)
  double* ds = inrec + N;                   // N: broadcast copy of d
  double* rot = inrec + 2 * N;              // 3N: (cc, ss, xny) per rotation
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

  const bool act = lane < N;
  const int hL = (lane >= NH) ? 1 : 0;
  const int kL = act ? lane - hL * NH : 0;
  const double dt = P.dt_mpc, b0 = dt * dt / 2, b1 = dt;
  const double thmax = P.theta_lim, thmin = -P.theta_lim;
  const double j_ini = P.j_ini, tq = P.torque_lim / P.j_ini;
  const double inf = CUDART_INF;

  // list mode: the instances body_split_kernel handed over (all of them if its list overflowed)
  int nB = P.B;
  bool listed = false;
  if (P.flist) {
    const int cnt = *reinterpret_cast<volatile const int*>(P.flist_count);
    if (cnt <= P.flist_cap) { nB = cnt; listed = true; }
  }

  for (;;) {
    int b = 0;
    if (lane == 0) {
      b = atomicAdd(P.sched, 1);
      if (listed && b < nB) b = P.flist[b]; else if (listed) b = 0x7fffffff;
    }
    b = __shfl_sync(FULL_MASK, b, 0);
    if (b >= P.B) break;

    // ---- stage the input record (one TMA bulk copy) ----
    fence_proxy_async();     // the scratch that overlays the record was written through the generic proxy
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
    const double xw = act ? inrec[36 + lane] : 0.0;          // warm start entry (kept if G is not PD / gated)
    const double* refs = inrec + 36 + N;

    int status = -1, iq = 0, bjx1 = 0, bjx2 = 0;
    int it_outer = 0, it_add = 0, it_drop = 0, it_degen = 0, it_l2a = 0;
    unsigned flops = 0;
    double f_value = 0.0;
    int A = 0;   // working-set slot owned by this lane

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
      if (act) outrec[18 + lane] = xw;
      if (lane == 0) outrec[18 + N] = 0.0;
    } else {
      // ---- phase indices (cpp:406-417): first table entry the time has not reached ----
      {
        double txl = (lane < 27) ? inrec[lane] : inf;
        unsigned g1 = __ballot_sync(FULL_MASK, !((i + 1) * dt >= txl));
        unsigned g2 = __ballot_sync(FULL_MASK, !((i + NH) * dt >= txl));
        bjx1 = __ffs(g1) - 1;    // = (j - 1) + 1 of the reference's while loop
        bjx2 = __ffs(g2) - 1;
      }
      const int t_yu = (i + 1) % P.nstepx;
      const bool left = (bjx1 < 2) || (bjx1 % 2 == 0);
      const bool sw = (bjx1 >= 2) && !((t_yu + NH - 1) < P.nstepx);
      const int t_yu_k = (t_yu + NH) - P.nstepx;

      // ---- condensation (cpp:427-526): lane L owns horizon step kL of half hL ----
      const double my0 = hL ? thy0 : thx0, my1 = hL ? thy1 : thx1;
      double g0 = 0.0, pk = 0.0, pth = 0.0;
      double zx_o = 0.0, zy_o = 0.0, caz_o = 0.0;            // lanes 0..2: outputs need steps 0..2
      const double bs_p = (lane < 2) ? inrec[32 + 2 * lane] : 0.0, bs_v = (lane < 2) ? inrec[33 + 2 * lane] : 0.0;
      if (lane < 3) { zx_o = refs[lane]; zy_o = refs[NH + lane]; caz_o = refs[8 * NH + lane]; }
      if (act) {
        const int k = kL;
        const bool other = sw && (k >= NH - t_yu_k);
        const bool use_l = left ? !other : other;
        const double cop = hL ? (use_l ? refs[6 * NH + k] : refs[4 * NH + k])      // half 1 (pitch accel) uses det_px
                              : (use_l ? refs[7 * NH + k] : refs[5 * NH + k]);     // half 0 (roll accel) uses det_py
        const double det = (hL ? refs[k] : refs[NH + k]) - cop;
        pth = j_ini / (P.mass * (refs[8 * NH + k] + P.g));
        pk = fma(pps[k], my0, pps[NH + k] * my1);
        const double t1 = fma(m1[k], my0, m1[NH + k] * my1), t2 = fma(m2[k], my0, m2[NH + k] * my1);
        const double* bref = refs + (2 + hL) * NH;
        double t3 = 0.0;
#pragma unroll
        for (int j = 0; j < NH; j++) t3 = fma(s2[j * NH + k], bref[j], t3);
        g0 = ((t1 + t2) - t3) + (P.gama * (hL ? -pth : pth)) * det;
      }
      __syncwarp();   // every input field is in registers: the record area becomes scratch

      // ---- Hessian block, Cholesky in registers (lanes own rows), pivot by rsqrt ----
      double Lrow[NH];
      double tr = 0.0;
#pragma unroll
      for (int j = 0; j < NH; j++) {
        double v = (lane < NH) ? gc0[j * NH + lane] : 0.0;
        if (j == lane) { v = v + P.gama / 2 * (pth * pth); tr = 2 * v; }
        Lrow[j] = 2 * v;
      }
      tr = warp_sum(tr);
      const double c1 = 2 * tr;
      flops = (unsigned)gi_flops_setup(N, 0);
      bool pd = true;
      double linv = 0.0;   // lane k: 1 / L(k,k)
#pragma unroll
      for (int k = 0; k < NH; k++) {
        double v = Lrow[k];
#pragma unroll
        for (int j = 0; j < k; j++) v = fma(-Lrow[j], bcast(Lrow[j], k), v);
        const double piv = bcast(v, k);
        if (!(piv > 0.0) && !(piv != piv)) { pd = false; break; }   // <= 0: not PD (NaN passes, as in the reference)
        const double rs = rsqrt(piv);
        Lrow[k] = (lane >= k) ? v * rs : 0.0;
        if (lane == k) linv = rs;
      }

      double x = xw, xold = 0.0, z = 0.0, u = 0.0, uold = 0.0, r = 0.0, rinv = 0.0;
      int Aold = 0;
      if (!pd) {
        status = ST_NOT_PD;   // x keeps the warm start, exactly as the reference
        f_value = inf;
      } else {
        // L and 1/diag to scratch (R area), then J = L^-T: every lane builds column kL
        double* Ls = Rp;            // L(k,i) = Ls[i*NH + k]
        double* linvs = Rp + NH * NH;
        if (lane < NH) {
#pragma unroll
          for (int j = 0; j < NH; j++) Ls[j * NH + lane] = Lrow[j];
          linvs[lane] = linv;
        }
        if (act) xs[lane] = g0;
        __syncwarp();
        double y[NH];
        double c2 = 0.0;
#pragma unroll
        for (int ii = NH - 1; ii >= 0; ii--) {
          double t = 0.0;
#pragma unroll
          for (int k = ii + 1; k < NH; k++) t = fma(Ls[ii * NH + k], y[k], t);
          const double li = linvs[ii];
          y[ii] = (ii == kL) ? li : ((ii < kL) ? -t * li : 0.0);
          c2 += li;
        }
        c2 = 2 * c2;
        // d0 = J' g0 (column owner has the column in registers), columns to shared memory
        double d0 = 0.0;
#pragma unroll
        for (int ii = 0; ii < NH; ii++) d0 = fma(y[ii], xs[hL * NH + ii], d0);
        __syncwarp();
        if (act) {
          double* col = J + lane * LD;
#pragma unroll
          for (int ii = 0; ii < NH; ii++) { col[hL * NH + ii] = y[ii]; col[(1 - hL) * NH + ii] = 0.0; }
          ds[lane] = d0;
        }
        __syncwarp();
        // x = -J d0, f = g0.x / 2
        {
          double acc = 0.0;
          if (act) {
            double c0 = 0.0, c1 = 0.0, c2_ = 0.0, c3 = 0.0;
#pragma unroll
            for (int j = 0; j < N; j += 4) {
              c0 = fma(J[j * LD + lane], ds[j], c0);
              if (j + 1 < N) c1 = fma(J[(j + 1) * LD + lane], ds[j + 1], c1);
              if (j + 2 < N) c2_ = fma(J[(j + 2) * LD + lane], ds[j + 2], c2_);
              if (j + 3 < N) c3 = fma(J[(j + 3) * LD + lane], ds[j + 3], c3);
            }
            acc = (c0 + c1) + (c2_ + c3);
          }
          x = -acc;
          double f = act ? g0 * x : 0.0;
          f_value = 0.5 * warp_sum(f);
        }
        __syncwarp();

        // ================= Goldfarb-Idnani iteration (EiQuadProg.cpp:282-490) =================
        const double tol = M * EPS_D * c1 * c2 * 100.0;
        const int cap = P.cap_scale * (N + M) + 50;
        double R_norm = 1.0;
        status = ST_OK;
        unsigned inA = 0u, excl = 0u, inAold = 0u;   // bit s <-> this lane's constraint slot s
        double s0 = 0, s1 = 0, s2v = 0, s3 = 0;      // slacks: angle up, angle low, torque up, torque low
        // constraint index of slot s of this lane: blk = ((s>>1)<<2) | (hL<<1) | (s&1)
        const int cbase = (hL << 1) * NH + kL;
        enum { PH_L1, PH_L2, PH_L2A };
        int ph = PH_L1, ip = 0, passes = 0;
        double ss = 0.0, sip = 0.0, npL = 0.0;
        int ip_blk = 0, ip_k = 0, ip_h = 0;
        double ip_sgn = 1.0;

        auto slack_angle = [&](double& up, double& low) {
          double v0 = 0.0, v1 = 0.0;       // two chains: halves the dependent-FMA latency
#pragma unroll
          for (int j = 0; j < NH; j += 2) {
            if (j <= kL) v0 = fma(ppu[j * NH + kL], xs[hL * NH + j], v0);
            if (j + 1 < NH && j + 1 <= kL) v1 = fma(ppu[(j + 1) * NH + kL], xs[hL * NH + j + 1], v1);
          }
          const double v = v0 + v1;
          up = (thmax - pk) - v;
          low = v + (thmax + pk);
        };

        for (;;) {
          if (ph == PH_L1) {
            it_outer++;
            flops += 2u * N * M;
            if (act) xs[lane] = x;
            __syncwarp();
            double psi = 0.0;
            if (act) {
              slack_angle(s0, s1);
              s2v = fma(-j_ini, x, tq);
              s3 = fma(j_ini, x, tq);
              psi = (fmin(0.0, s0) + fmin(0.0, s1)) + (fmin(0.0, s2v) + fmin(0.0, s3));
            }
            psi = warp_sum(psi);
            excl = 0u;
            ss = 0.0; ip = 0;
            if (fabs(psi) <= tol) break;
            uold = u; Aold = A; xold = x; inAold = inA;
            ph = PH_L2;
          }
          if (ph == PH_L2) {
            // most negative eligible slack, lowest constraint index among equals (cpp:322-342)
            double bv = ss; int bi = 0x7fffffff;
            if (act) {
              const unsigned blocked = inA | excl;
              // slots in ascending constraint index: 0 (blk hL*2), 1, 2 (blk 4+hL*2), 3
              if (!(blocked & 1u) && s0 < bv) { bv = s0; bi = cbase; }
              if (!(blocked & 2u) && s1 < bv) { bv = s1; bi = cbase + NH; }
              if (!(blocked & 4u) && s2v < bv) { bv = s2v; bi = cbase + 4 * NH; }
              if (!(blocked & 8u) && s3 < bv) { bv = s3; bi = cbase + 5 * NH; }
            }
            warp_argmin_redux(bv, bi);
            if (bv < ss) { ss = bv; ip = bi; }
            if (ss >= 0.0) break;
            sip = ss;
            ip_blk = ip / NH; ip_k = ip - ip_blk * NH;
            ip_h = (ip_blk >> 1) & 1;
            ip_sgn = (ip_blk & 1) ? 1.0 : -1.0;
            // own entry of n+ = CI(:, ip)
            npL = 0.0;
            if (act && hL == ip_h) {
              if (ip_blk < 4) { if (kL <= ip_k) npL = ip_sgn * ppu[kL * NH + ip_k]; }
              else if (kL == ip_k) npL = ip_sgn * j_ini;
            }
            if (lane == iq) { u = 0.0; A = ip; }
            ph = PH_L2A;
          }
          // ---- step 2a (cpp:349-386) ----
          if (++passes > cap) { status = ST_ITER_CAP; break; }
          it_l2a++;
          flops += 2u * N * N + 2u * N * (N - iq) + (unsigned)(iq * iq) + 4u * N + 2u * iq;
          // d = J' n+ : lane owns column `lane`; n+ comes from the model table
          double d = 0.0;
          if (act) {
            const double* col = J + lane * LD + ip_h * NH;
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
            ds[lane] = d;
          }
          __syncwarp();
          // z = J[:, iq:] d[iq:]
          z = 0.0;
          if (act) {
            // run-time start (late passes of a long solve have few free columns left), two chains
            double z0 = 0.0, z1 = 0.0;
            int j = iq;
            if ((N - j) & 1) { z0 = J[j * LD + lane] * ds[j]; j++; }
#pragma unroll 2
            for (; j < N; j += 2) {
              z0 = fma(J[j * LD + lane], ds[j], z0);
              z1 = fma(J[(j + 1) * LD + lane], ds[j + 1], z1);
            }
            z = z0 + z1;
          }
          // r = R^-1 d[0:iq)  (column-oriented back substitution, stored reciprocals)
          r = (lane < iq) ? d : 0.0;
          for (int c = iq - 1; c >= 0; c--) {
            const double rc = bcast(r * rinv, c);
            if (lane == c) r = rc;
            else if (lane < c) r = fma(-rc, Rp[c * (c + 3) / 2 + lane], r);
          }
          // step lengths: t1 over the working set, t2 along z.  ONE division instruction serves the
          // ratio test (lanes < iq: u/r), t2 (lane 31: -s_ip / z.n+) and the Householder scale of a
          // possible add (lane 30: 1 / (|d2| (|d2| + |d_iq|))): FP64 division is a ~40-instruction
          // dependent sequence, the longest scalar latency of a pass.
          double zz = z * z, zn = z * npL, dd = (act && lane >= iq) ? d * d : 0.0;
          warp_sum3(zz, zn, dd);
          const double inrm = (dd > 0.0) ? rsqrt(dd) : 0.0;      // 1 / |d2|
          const double nrm = dd * inrm;                           // |d2|
          const double diq = bcast(d, iq);
          double num = u, den = r;
          if (lane == 31) { num = -sip; den = zn; }
          if (lane == 30) { num = 1.0; den = nrm * (nrm + fabs(diq)); }
          const double quo = num / den;
          double t1 = inf; int kmin = 0x7fffffff;
          if (lane < iq && r > 0.0) { t1 = quo; kmin = lane; }
          warp_argmin_redux(t1, kmin);
          const int l = (kmin != 0x7fffffff && t1 < inf) ? __shfl_sync(FULL_MASK, A, kmin) : 0;
          const double t2 = (fabs(zz) > EPS_D) ? bcast(quo, 31) : inf;
          const double tau = bcast(quo, 30);
          const double t = fmin(t1, t2);
          if (t >= inf) { status = ST_INFEASIBLE; f_value = inf; break; }        // case (i)
          int qq = 0;
          bool do_drop = false;
          if (t2 >= inf) {                                                        // case (ii): dual step
            if (lane < iq) u = fma(-t, r, u);
            if (lane == iq) u += t;
            do_drop = true;
          } else {
            // case (iii): step in primal and dual space
            const double uiq = bcast(u, iq);
            x = fma(t, z, x);
            f_value += t * zn * (0.5 * t + uiq);
            if (lane < iq) u = fma(-t, r, u);
            if (lane == iq) u = uiq + t;
            if (t == t2) {
              // ---- add_constraint (cpp:30-93) ----
              // The reference zeroes d[iq+1:] with a chain of n-iq-1 Givens rotations of J's trailing
              // columns.  Any orthogonal map that sends d2 = d[iq:] to a multiple of e_0 spans the same
              // null-space basis, so here ONE Householder reflection H = I - tau v v' does it,
              // v = d2 + sigma e_0, sigma = sign(d_iq) |d2|:  J2 <- J2 H = J2 - tau (J2 v) v' and
              // J2 v = J2 d2 + sigma J(:,iq) = z + sigma J(:,iq) -- z is already in registers, so the
              // update is a single dependency-free pass over the trailing columns.
              flops += 6u * N * (unsigned)(N - iq - 1 > 0 ? N - iq - 1 : 0);
              // |d2| = nrm, 1/|d2| = inrm, tau: computed with the step lengths above
              const double sigma = (diq < 0.0) ? -nrm : nrm;
              if (nrm != 0.0 && act) {
                const double sw = tau * fma(sigma, J[iq * LD + lane], z);   // tau * (J2 v)_L
                const double viq = diq + sigma;
                J[iq * LD + lane] = fma(-sw, viq, J[iq * LD + lane]);
#pragma unroll 4
                for (int j = iq + 1; j < N; j++) J[j * LD + lane] = fma(-sw, ds[j], J[j * LD + lane]);
              }
              // H d2 = -sigma e_0
              if (lane == iq) d = (nrm != 0.0) ? -sigma : d;
              if (act && lane > iq) d = 0.0;
              // new column of R, its reciprocal diagonal, degeneracy test
              if (lane <= iq) Rp[iq * (iq + 3) / 2 + lane] = d;
              if (lane == iq) rinv = (nrm != 0.0) ? ((diq < 0.0) ? inrm : -inrm) : 1.0 / d;   // 1 / (-sigma)
              const double dq = (nrm != 0.0) ? -sigma : diq;           // new R(iq,iq), warp-uniform
              iq++;
              __syncwarp();
              if (fabs(dq) <= EPS_D * R_norm) {
                // degenerate (cpp:444-462): exclude ip, remove it again, restore the state saved at step 1
                it_degen++;
                // owner lane of ip marks it excluded
                {
                  const int own = ip_h * NH + ip_k, slot = ((ip_blk >> 2) << 1) | (ip_blk & 1);
                  if (lane == own) excl |= 1u << slot;
                }
                // delete_constraint(ip): ip sits in the last slot, no rotation needed
                iq--;
                if (lane == iq) { A = 0; u = 0.0; }
                if (lane < iq) { A = Aold; u = uold; }
                inA = inAold;
                x = xold;
                if (act) xs[lane] = x;
                __syncwarp();
                ph = PH_L2;
                continue;
              }
              R_norm = fmax(R_norm, fabs(dq));
              it_add++;
              {
                const int own = ip_h * NH + ip_k, slot = ((ip_blk >> 2) << 1) | (ip_blk & 1);
                if (lane == own) inA |= 1u << slot;
              }
              ph = PH_L1;
              continue;
            }
            do_drop = true;   // partial step: drop l, recompute s(ip), stay in 2a (cpp:477-490)
          }
          if (do_drop) {
            // ---- delete_constraint(l) (cpp:95-170) ----
            {
              const int lblk = l / NH, lk = l - lblk * NH;
              const int own = ((lblk >> 1) & 1) * NH + lk, slot = ((lblk >> 2) << 1) | (lblk & 1);
              if (lane == own) inA &= ~(1u << slot);
            }
            const unsigned hit = __ballot_sync(FULL_MASK, lane < iq && A == l);
            if (!hit) { status = ST_ITER_CAP; break; }   // l not in the working set: UB in the reference
            qq = __ffs(hit) - 1;
            {
              const int An = __shfl_down_sync(FULL_MASK, A, 1);
              const double un = __shfl_down_sync(FULL_MASK, u, 1);
              if (lane >= qq && lane < iq) { A = An; u = un; }
              if (lane == iq) { A = 0; u = 0.0; }
            }
            // R columns shift left by one (column c+1 -> c, rows 0..c+1)
            for (int c = qq; c < iq - 1; c++) {
              const double v = (lane <= c + 1) ? Rp[(c + 1) * (c + 4) / 2 + lane] : 0.0;
              __syncwarp();
              if (lane <= c + 1) Rp[c * (c + 3) / 2 + lane] = v;
            }
            iq--;
            __syncwarp();
            it_drop++;
            flops += 3u * (unsigned)((iq - qq) * (iq - qq)) + 6u * N * (unsigned)(iq - qq);
            for (int j = qq; j < iq; j++) {
              double cc = Rp[j * (j + 3) / 2 + j], sn = Rp[j * (j + 3) / 2 + j + 1];
              const double h = gi_hypot(cc, sn);
              __syncwarp();
              if (h == 0.0) continue;
              cc = cc / h; sn = sn / h;
              const double diag = (cc < 0.0) ? -h : h;
              if (lane == 0) { Rp[j * (j + 3) / 2 + j + 1] = 0.0; Rp[j * (j + 3) / 2 + j] = diag; }
              if (cc < 0.0) { cc = -cc; sn = -sn; }
              const double xny = sn / (1.0 + cc);
              if (lane > j && lane < iq) {
                double* cp = Rp + lane * (lane + 3) / 2;
                const double t1j = cp[j], t2j = cp[j + 1];
                const double a = fma(t2j, sn, t1j * cc);
                cp[j] = a;
                cp[j + 1] = fma(xny, t1j + a, -t2j);
              }
              if (act) {
                const double t1j = J[j * LD + lane], t2j = J[(j + 1) * LD + lane];
                const double a = fma(t2j, sn, t1j * cc);
                J[j * LD + lane] = a;
                J[(j + 1) * LD + lane] = fma(xny, a + t1j, -t2j);
              }
              __syncwarp();
            }
            if (lane >= qq && lane < iq) rinv = 1.0 / Rp[lane * (lane + 3) / 2 + lane];
            if (t2 >= inf) continue;   // dual step: back to 2a with the same ip
            // partial step: recompute the slack of ip at the new x
            if (act) xs[lane] = x;
            __syncwarp();
            {
              const int own = ip_h * NH + ip_k;
              double sv = 0.0;
              if (lane == own) {
                if (ip_blk < 4) { double up, low; slack_angle(up, low); sv = (ip_blk & 1) ? low : up; }
                else sv = (ip_blk & 1) ? fma(j_ini, x, tq) : fma(-j_ini, x, tq);
              }
              sip = bcast(sv, own);
            }
          }
        }
      }

      // ---- first control: fallback / clamp (cpp:567-625) ----
      bool has_nan = __any_sync(FULL_MASK, act && (x != x));
      if (has_nan && (status == ST_OK || status == ST_EQ_DEP)) status = ST_NAN;
      double ax0 = bcast(x, 0), ay0 = bcast(x, NH);
      const double arow_x = thx0 + dt * thx1, arow_y = thy0 + dt * thy1;
      if (has_nan) {
        ax0 = (thx0 - arow_x) / b0;
        ay0 = (thy0 - arow_y) / b0;
      } else {
        const double nx0 = arow_x + b0 * ax0;
        if (nx0 > thmax) ax0 = (thmax - arow_x) / b0;
        else if (nx0 < thmin) ax0 = (thmin - arow_x) / b0;
        const double ny0 = arow_y + b0 * ay0;
        if (ny0 > thmax) ay0 = (thmax - arow_y) / b0;
        else if (ny0 < thmin) ay0 = (thmin - arow_y) / b0;
      }
      if (lane == 0) x = ax0;
      if (lane == NH) x = ay0;
      // ---- roll-out (cpp:629-655): lane 0 integrates roll, lane 1 pitch; out14 needs steps 0..2 ----
      const double xa1 = bcast(x, 1), xa2 = bcast(x, 2), ya1 = bcast(x, NH + 1), ya2 = bcast(x, NH + 2);
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
        const double xacc = x;                               // lane jj: roll accel at step jj
        const double yacc = bcast(x, (NH + lane) & 31);      // pitch accel at step jj
        if (lane < 3) {
          const double den = P.mass * (P.g + caz_o);
          const int o = (lane == 0) ? 4 : (lane == 1 ? 8 : 12);
          outrec[o] = zx_o - j_ini * yacc / den;
          outrec[o + 1] = zy_o + j_ini * xacc / den;
        }
      }
      if (act) outrec[18 + lane] = x;
      if (lane == 0) outrec[18 + N] = f_value;
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
        dg[0] = status; dg[1] = iq;
        dg[2] = it_outer; dg[3] = it_add; dg[4] = it_drop; dg[5] = it_degen;
        dg[6] = bjx1; dg[7] = bjx2; dg[8] = it_l2a; dg[9] = (int)flops;
      }
      if (act) dg[10 + lane] = (lane < iq) ? A : -1;
    }
    if (lane == 0) tma_store_wait_read();
    __syncwarp();
  }
  if (lane == 0) {
    tma_store_wait_all();
    // last warp out resets the instance counter for the next launch
    const int done = atomicAdd(P.sched + 1, 1);
    if (done == (int)(gridDim.x * WPC) - 1) {
      P.sched[0] = 0; P.sched[1] = 0;
      if (P.flist) { P.flist_count[1] += P.flist_count[0]; P.flist_count[0] = 0; }   // [1]: running total, read by go1mpc_body_handover_total
      __threadfence();
    }
  }
}

template <int NH>
static size_t fast_smem(int wpc) {
  using D = FastDims<NH>;
  return (size_t)(D::TAB + wpc * D::WD) * sizeof(double) + (size_t)(wpc + 1) * sizeof(uint64_t);
}

template <int NH>
static cudaError_t fast_launch_nh(const BodyKParams& P, int sms, cudaStream_t st, int* grid_out) {
  constexpr int WPC = 4;
  using D = FastDims<NH>;
  if (P.in_stride != D::IN || P.out_stride != D::OUT || P.tab_doubles != D::TAB) return cudaErrorInvalidValue;
  const size_t smem = fast_smem<NH>(WPC);
  int dev_ = 0;            // function attributes and occupancy are per device
  cudaGetDevice(&dev_);
  static std::mutex cache_mu;         // handles on several host threads may launch concurrently
  static int occ_cache_[64] = {};
  int occ_cache;
  {
    std::lock_guard<std::mutex> lk(cache_mu);
    if (occ_cache_[dev_ & 63] == 0) {
      cudaError_t e = cudaFuncSetAttribute(body_fast_kernel<NH, WPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      int o = 0;
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, body_fast_kernel<NH, WPC>, WPC * 32, smem);
      if (e != cudaSuccess) return e;
      if (o < 1) return cudaErrorLaunchOutOfResources;
      occ_cache_[dev_ & 63] = o;
    }
    occ_cache = occ_cache_[dev_ & 63];
  }
  int grid = (P.B + WPC - 1) / WPC;
  if (grid > sms * occ_cache) grid = sms * occ_cache;
  if (P.flist && grid > sms) grid = sms;     // list mode: the list is short (normally empty)
  if (grid_out) *grid_out = grid;
  body_fast_kernel<NH, WPC><<<grid, WPC * 32, smem, st>>>(P);
  return cudaGetLastError();
}

bool body_fast_supported(int nh) { return nh == 4 || nh == 10; }

cudaError_t body_fast_launch(BodyKParams P, int sms, cudaStream_t st) {
  switch (P.nh) {
    case 4: return fast_launch_nh<4>(P, sms, st, nullptr);
    case 10: return fast_launch_nh<10>(P, sms, st, nullptr);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace go1
