// body_duo.cu -- body-inclination MPC tick for ANY horizon 3..40 with the block structure of its QP exploited:
// one warp per instance, the roll and the pitch half of the solver state side by side in the warp's shared memory.
//
// Same contract as body_mpc.cu (replaces PRMPCClass::body_theta_mpc, RT/src/FastMPC/PRMPCClass.cpp:379-714,
// solve_body_rotation / Solve :799-849, Indexfind :716-738; QP = Eigen::QP::solve_quadprog2,
// RT/src/utils/EiQuadProg/EiQuadProg.cpp:172-491).  body_mpc.cu treats the QP as a dense n = 2 nh problem: J and R are
// (2 nh)^2 each -- 104 KB per warp at nh = 40, two warps per SM -- and every pass pays for 2 nh columns.  But
// G = blockdiag(H, H) and every constraint column has its support in one half, so J = L^-T stays block diagonal under
// the solver's orthogonal updates and the cross-half entries of d, z, r, R are exact zeros (the argument is spelled
// out in body_tri.cu / DESIGN.md 3.1): the reference's 2 nh-variable solve is an interleaving of two nh-variable
// solves that share only the selection of the next constraint (most negative slack over BOTH halves, lowest index on
// ties), the stopping test on the summed infeasibility, R_norm and the iteration counters.  This kernel runs exactly
// that interleaving in one warp -- no log, no replay:
//   * per half: J (nh x nh), R, z, d, r, the duals and the half's working set; per pass only the half of the entering
//     constraint is touched (a quarter of the dense pass's flops, a quarter of its shared memory);
//   * shared: x, the slacks s (only the half that moved since the last step 1 is re-evaluated), the ORDERED global
//     working set (what the reference's A holds), psi per half, R_norm, the counters and the algorithmic flop count
//     of the dense algorithm (what bench.py's roofline counts);
//   * the solver steps themselves are gi_warp.cuh's (Householder add, Givens drop), called on the half's workspace.
// One corner is handed to body_mpc.cu's dense kernel through the hand-over list instead of being guessed: a degenerate
// add after a drop inside the same outer iteration (the reference then restores A / u positionally from a snapshot of
// another length; which garbage it continues with depends on the global column order).
#include <cuda_runtime.h>
#include <stdint.h>
#include <mutex>
#include "gi_warp.cuh"
#include "tma.cuh"
#include "kernels.h"

namespace go1 {

namespace {

__device__ __forceinline__ int duo_half_of(int c, int nh) { return ((c / nh) >> 1) & 1; }

struct DuoShared {
  int nh;
  const double* ppu;    // CTA-shared, nh x nh column-major (lower triangular)
  const double* ppsx;   // nh: Pps * thetaxk
  const double* ppsy;
  double j_ini, thmax, tq;
  double* X;            // 2 nh
  double* S;            // 8 nh

  __device__ __forceinline__ double s_of(int c) const {
    const int blk = c / nh, k = c - blk * nh;
    if (blk < 4) {
      const int half = blk >> 1, low = blk & 1;
      const double* xx = X + half * nh;
      double v = 0.0;
      for (int j = 0; j <= k; j++) v = fma(ppu[j * nh + k], xx[j], v);
      const double pk = (half ? ppsy : ppsx)[k];
      return low ? (v + (thmax + pk)) : ((thmax - pk) - v);
    }
    const int b2 = blk - 4, half = b2 >> 1, low = b2 & 1;
    const double xv = X[half * nh + k];
    return low ? fma(j_ini, xv, tq) : fma(-j_ini, xv, tq);
  }
  // slacks of one half (its 4 nh constraints), returns the lane's part of psi
  __device__ __forceinline__ double eval_half(int h, int lane) const {
    double psi = 0.0;
    const double* xx = X + h * nh;
    const double* ppk = h ? ppsy : ppsx;
    for (int k = lane; k < nh; k += 32) {
      // one triangular product per horizon step serves the upper and the lower angle bound (same value as s_of)
      double v = 0.0;
      for (int j = 0; j <= k; j++) v = fma(ppu[j * nh + k], xx[j], v);
      const double pk = ppk[k], xv = xx[k];
      const double s0 = (thmax - pk) - v, s1 = v + (thmax + pk);
      const double s2 = fma(-j_ini, xv, tq), s3 = fma(j_ini, xv, tq);
      S[(2 * h) * nh + k] = s0; S[(2 * h + 1) * nh + k] = s1;
      S[(4 + 2 * h) * nh + k] = s2; S[(5 + 2 * h) * nh + k] = s3;
      psi += (fmin(0.0, s0) + fmin(0.0, s1)) + (fmin(0.0, s2) + fmin(0.0, s3));
    }
    return psi;
  }
  // n+ of constraint ip in its half's coordinates: np[0 .. nh), non-zero range [klo, khi)
  __device__ __forceinline__ void load_np(double* np, int ip, int lane, int& klo, int& khi) const {
    const int blk = ip / nh, k = ip - blk * nh;
    int low, lo, hi;
    if (blk < 4) { low = blk & 1; lo = 0; hi = k + 1; }
    else { low = (blk - 4) & 1; lo = k; hi = k + 1; }
    for (int j = lane; j < nh; j += 32) {
      double v = 0.0;
      if (j >= lo && j < hi) {
        v = (blk < 4) ? ppu[j * nh + k] : j_ini;
        if (!low) v = -v;
      }
      np[j] = v;
    }
    klo = lo; khi = hi;
    __syncwarp();
  }
};

// per-half workspace carved from the warp's slice: J | R | z d np r u uold (nh + 2 each) | A Aold (ints)
__host__ __device__ inline int duo_half_doubles(int nh) {
  const int ld = gi_ld(nh), v = nh + 2;
  int d = 2 * nh * ld + 6 * v;
  d += (2 * v + 1) / 2;
  return (d + 1) & ~1;
}
__device__ inline void duo_carve(GiWs& w, double* base, int nh, double* x, double* xold, double* s) {
  const int ld = gi_ld(nh), v = nh + 2;
  w.n = nh; w.p = 0; w.m = 0; w.ld = ld; w.ms = 0;
  w.J = base; base += nh * ld;
  w.R = base; base += nh * ld;
  w.z = base; base += v; w.d = base; base += v; w.np = base; base += v;
  w.r = base; base += v; w.u = base; base += v; w.uold = base; base += v;
  w.A = reinterpret_cast<int*>(base);
  w.Aold = w.A + v;
  w.x = x; w.xold = xold; w.s = s; w.rot = nullptr;
}

// The interleaved main loop (the structure of gi_loop, gi_warp.cuh).  Requires per half: J = L^-T, R = 0; X = the
// unconstrained minimiser, res.f its cost.  G receives the ordered working set.  Returns false when the instance must go
// to the dense kernel (see the header).
__device__ inline bool duo_loop(GiWs& w0, GiWs& w1, const DuoShared& sh, double* XOLD, int* G, double c1, double c2, int cap,
                                GiResult& res, int lane) {
  const int nh = sh.nh, n = 2 * nh, m = 12 * nh, ms = 8 * nh;
  const double inf = CUDART_INF;
  double R_norm = 1.0, f_value = res.f;
  int iqh[2] = {0, 0};
  int status = ST_OK;
  int it_outer = 0, it_add = 0, it_drop = 0, it_degen = 0, it_l2a = 0;
  unsigned long long flops = res.flops;
  unsigned inA = 0u, excl = 0u;    // bit t <-> constraint lane + 32 t
  bool ok = true;
  for (int t = lane; t < nh + 2; t += 32) {
    w0.u[t] = 0.0; w0.uold[t] = 0.0; w0.A[t] = 0; w0.Aold[t] = 0; w0.r[t] = 0.0;
    w1.u[t] = 0.0; w1.uold[t] = 0.0; w1.A[t] = 0; w1.Aold[t] = 0; w1.r[t] = 0.0;
  }
  __syncwarp();
  enum { PH_L1, PH_L2, PH_L2A };
  int phase = PH_L1, ip = 0, l = 0, passes = 0, klo = 0, khi = nh, hh = 0, drops_outer = 0;
  bool dirty0 = true, dirty1 = true;
  double psi0 = 0.0, psi1 = 0.0, ss = 0.0;
  for (;;) {
    const int iq = iqh[0] + iqh[1];
    if (phase == PH_L1) {
      // EiQuadProg.cpp:282-320
      it_outer++;
      flops += 2ull * n * m;
      inA = 0u;
      for (int i = 0; i < iq; i++) { const int c = G[i]; if ((c & 31) == lane) inA |= 1u << (c >> 5); }
      excl = 0u;
      if (dirty0) { psi0 = warp_sum(sh.eval_half(0, lane)); dirty0 = false; }
      if (dirty1) { psi1 = warp_sum(sh.eval_half(1, lane)); dirty1 = false; }
      const double psi = psi0 + psi1;
      ss = 0.0; ip = 0; drops_outer = 0;
      __syncwarp();
      if (fabs(psi) <= m * EPS_D * c1 * c2 * 100.0) break;
      for (int t = lane; t < iqh[0]; t += 32) { w0.uold[t] = w0.u[t]; w0.Aold[t] = w0.A[t]; }
      for (int t = lane; t < iqh[1]; t += 32) { w1.uold[t] = w1.u[t]; w1.Aold[t] = w1.A[t]; }
      for (int k = lane; k < n; k += 32) XOLD[k] = sh.X[k];
      __syncwarp();
      phase = PH_L2;
    }
    GiWs& w = hh ? w1 : w0;      // rebound below once ip is chosen
    if (phase == PH_L2) {
      // EiQuadProg.cpp:322-342: most negative eligible s over both halves, first index wins
      double bv = ss; int bi = 0x7fffffff;
      for (int c = lane, t = 0; c < ms; c += 32, t++) {
        const double sv = sh.S[c];
        if (sv < bv && !((inA >> t) & 1u) && !((excl >> t) & 1u)) { bv = sv; bi = c; }
      }
      warp_argmin_redux(bv, bi);
      if (bv < ss) { ss = bv; ip = bi; }
      if (ss >= 0.0) break;
      hh = duo_half_of(ip, nh);
      GiWs& wn = hh ? w1 : w0;
      sh.load_np(wn.np, ip, lane, klo, khi);
      if (lane == 0) { wn.u[iqh[hh]] = 0.0; wn.A[iqh[hh]] = ip; }
      __syncwarp();
      phase = PH_L2A;
      continue;                  // re-enter with w bound to the chosen half
    }
    // PH_L2A: EiQuadProg.cpp:349-490, on the half of ip
    int& iql = iqh[hh];
    if (++passes > cap) { status = ST_ITER_CAP; break; }
    it_l2a++;
    flops += 2ull * n * n + 2ull * n * (n - iq) + (unsigned long long)iq * iq + 4ull * n + 2ull * iq;
    gi_compute_d(w, klo, khi, lane);
    gi_update_z(w, iql, lane);
    gi_update_r(w, iql, lane);
    double t1 = inf; int kmin = 0x7fffffff;
    for (int k = lane; k < iql; k += 32) {
      const double rk = w.r[k];
      if (rk > 0.0) { const double tmp = w.u[k] / rk; if (tmp < t1) { t1 = tmp; kmin = k; } }
    }
    warp_argmin_redux(t1, kmin);
    l = (kmin != 0x7fffffff && t1 < inf) ? w.A[kmin] : 0;
    double zz = 0.0, zn = 0.0;
    for (int k = lane; k < nh; k += 32) { zz = fma(w.z[k], w.z[k], zz); zn = fma(w.z[k], w.np[k], zn); }
    zz = warp_sum(zz); zn = warp_sum(zn);
    const double t2 = (fabs(zz) > EPS_D) ? (-sh.S[ip] / zn) : inf;
    const double t = fmin(t1, t2);
    if (t >= inf) { status = ST_INFEASIBLE; f_value = inf; break; }            // case (i)
    // global position of l in the ordered working set (for the flop count and the removal)
    auto remove_from_G = [&](int name, int count) -> int {
      int q = -1;
      for (int base = 0; base < count; base += 32) {
        const int i = base + lane;
        const unsigned hit = __ballot_sync(FULL_MASK, i < count && G[i] == name);
        if (hit) { q = base + __ffs(hit) - 1; break; }
      }
      if (q < 0) return -1;
      for (int base = q; base < count - 1; base += 32) {
        const int i = base + lane;
        int a = 0;
        if (i < count - 1) a = G[i + 1];
        __syncwarp();
        if (i < count - 1) G[i] = a;
        __syncwarp();
      }
      return q;
    };
    if (t2 >= inf) {                                                            // case (ii): dual step
      for (int k = lane; k < iql; k += 32) w.u[k] = fma(-t, w.r[k], w.u[k]);
      if (lane == 0) w.u[iql] += t;
      if ((l & 31) == lane) inA &= ~(1u << (l >> 5));
      __syncwarp();
      int qq_local;
      if (!gi_delete_constraint(w, iql, l, lane, qq_local)) { status = ST_ITER_CAP; break; }
      const int qg = remove_from_G(l, iq);
      if (qg < 0) { status = ST_ITER_CAP; break; }
      it_drop++; drops_outer++;
      { const int iqn = iq - 1; flops += 3ull * (iqn - qg) * (iqn - qg) + 6ull * n * (iqn - qg); }
      continue;
    }
    // case (iii): step in primal and dual space
    const double uiq = w.u[iql];
    __syncwarp();
    for (int k = lane; k < nh; k += 32) w.x[k] = fma(t, w.z[k], w.x[k]);
    f_value += t * zn * (0.5 * t + uiq);
    for (int k = lane; k < iql; k += 32) w.u[k] = fma(-t, w.r[k], w.u[k]);
    if (lane == 0) w.u[iql] = uiq + t;
    if (hh) dirty1 = true; else dirty0 = true;
    __syncwarp();
    if (t == t2) {
      flops += 6ull * n * (n - iq - 1 > 0 ? n - iq - 1 : 0);
      if (!gi_add_constraint(w, iql, R_norm, lane)) {
        // EiQuadProg.cpp:444-462 degenerate: exclude ip, restore the state saved at step 1
        if (drops_outer > 0) { ok = false; break; }          // positional restore from a longer snapshot: dense kernel
        it_degen++;
        if ((ip & 31) == lane) excl |= 1u << (ip >> 5);
        int qq_local;
        if (!gi_delete_constraint(w, iql, ip, lane, qq_local)) { status = ST_ITER_CAP; break; }
        for (int t3 = lane; t3 < iqh[0]; t3 += 32) { w0.A[t3] = w0.Aold[t3]; w0.u[t3] = w0.uold[t3]; }
        for (int t3 = lane; t3 < iqh[1]; t3 += 32) { w1.A[t3] = w1.Aold[t3]; w1.u[t3] = w1.uold[t3]; }
        for (int k = lane; k < n; k += 32) sh.X[k] = XOLD[k];
        __syncwarp();
        inA = 0u;
        for (int i = 0; i < iq; i++) { const int c = G[i]; if ((c & 31) == lane) inA |= 1u << (c >> 5); }
        if (hh) dirty1 = false; else dirty0 = false;         // x is back where the slacks were evaluated
        phase = PH_L2;
        continue;
      }
      if (lane == 0) G[iq] = ip;
      __syncwarp();
      it_add++;
      if ((ip & 31) == lane) inA |= 1u << (ip >> 5);
      phase = PH_L1;
      continue;
    }
    // partial step: drop l, recompute s(ip), stay in 2a   (EiQuadProg.cpp:477-490)
    if ((l & 31) == lane) inA &= ~(1u << (l >> 5));
    {
      int qq_local;
      if (!gi_delete_constraint(w, iql, l, lane, qq_local)) { status = ST_ITER_CAP; break; }
      const int qg = remove_from_G(l, iq);
      if (qg < 0) { status = ST_ITER_CAP; break; }
      it_drop++; drops_outer++;
      const int iqn = iq - 1;
      flops += 3ull * (iqn - qg) * (iqn - qg) + 6ull * n * (iqn - qg);
      const double sv = sh.s_of(ip);
      if (lane == 0) sh.S[ip] = sv;
      __syncwarp();
    }
  }
  res.f = f_value; res.iq = iqh[0] + iqh[1]; res.status = status;
  res.it_outer = it_outer; res.it_add = it_add; res.it_drop = it_drop; res.it_degen = it_degen;
  res.it_l2a = it_l2a; res.flops = flops;
  return ok;
}

__device__ __forceinline__ int duo_indexfind(const double* tx, double goal) {
  int j = 0;
  while (j < 27 && goal >= tx[j]) j++;
  return j - 1;
}

constexpr int DUO_MAX_THREADS = 128;
#ifndef GO1_DUO_MINB
#define GO1_DUO_MINB 4
#endif

}  // namespace

// shared memory: [ppu nh^2 | m1 m2 pps 6 nh] per CTA, then per warp: two half workspaces | X XOLD (2 nh each) | S 8 nh |
// g0 2 nh | pth ppsx ppsy (nh each) | input record | output record | G (2 nh + 2 ints)
__host__ __device__ inline int duo_warp_doubles(int nh, int in_stride, int out_stride) {
  const int nhp = (nh + 1) & ~1;      // keeps the record buffers 16-byte aligned for odd horizons
  int d = 2 * duo_half_doubles(nh) + 4 * nh + 8 * nh + 2 * nh + 3 * nhp + in_stride + out_stride;
  d += (2 * nh + 2 + 1) / 2;
  return (d + 1) & ~1;
}
__host__ __device__ inline int duo_cta_doubles(int nh) { return (nh * nh + 6 * nh + 1) & ~1; }

__global__ void __launch_bounds__(DUO_MAX_THREADS, GO1_DUO_MINB) body_duo_kernel(BodyKParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* smem = reinterpret_cast<double*>(smem_raw);
  const int nh = P.nh, n = 2 * nh, m = 12 * nh;
  const int wpc = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // model table in HBM (api.cu: build_body_model): ppu | gc0 | s2 | m1 | m2 | pps
  const double* gtab = P.tab;
  const double* g_gc0 = gtab + nh * nh;
  const double* g_s2 = gtab + 2 * nh * nh;
  double* ppu = smem;
  double* m1 = smem + nh * nh;
  double* m2 = m1 + 2 * nh;
  double* pps = m2 + 2 * nh;
  for (int t = threadIdx.x; t < nh * nh; t += blockDim.x) ppu[t] = gtab[t];
  for (int t = threadIdx.x; t < 6 * nh; t += blockDim.x) m1[t] = gtab[3 * nh * nh + t];
  const int cta_d = duo_cta_doubles(nh);
  double* wbase = smem + cta_d + (size_t)warp * P.warp_doubles;
  double* X = wbase + 2 * duo_half_doubles(nh);
  double* XOLD = X + n;
  double* S = XOLD + n;
  double* g0 = S + 8 * nh;
  double* pth = g0 + n;
  const int nhp = (nh + 1) & ~1;
  double* ppsx = pth + nhp;
  double* ppsy = ppsx + nhp;
  double* inrec = ppsy + nhp;
  double* outrec = inrec + P.in_stride;
  int* G = reinterpret_cast<int*>(outrec + P.out_stride);
  GiWs w0, w1;
  duo_carve(w0, wbase, nh, X, XOLD, S);
  duo_carve(w1, wbase + duo_half_doubles(nh), nh, X + nh, XOLD + nh, S);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + cta_d + (size_t)wpc * P.warp_doubles);
  uint64_t* my_bar = bars + warp;
  if (threadIdx.x == 0) {
    for (int i = 0; i < wpc; i++) mbar_init(bars + i, 1);
    fence_mbar_init();
  }
  __syncthreads();
  uint32_t phase = 0;
  const double dt = P.dt_mpc;
  const double b0 = dt * dt / 2, b1 = dt;
  const double thmax = P.theta_lim, thmin = -P.theta_lim;
  const int ld = w0.ld;

  for (int b = blockIdx.x * wpc + warp; b < P.B; b += gridDim.x * wpc) {
    if (lane == 0) {
      mbar_expect_tx(my_bar, (uint32_t)(P.in_stride * sizeof(double)));
      tma_load_1d(inrec, P.in + (size_t)b * P.in_stride, (uint32_t)(P.in_stride * sizeof(double)), my_bar);
    }
    __syncwarp();
    mbar_wait(my_bar, phase);
    phase ^= 1u;

    const double* tx = inrec;
    const int tick = (int)inrec[27];
    const double* theta_in = inrec + 28;
    const double* bstate = inrec + 32;
    const double* xwarm = inrec + 36;
    const double* refs = inrec + 36 + 2 * nh;
    const double *zx = refs, *zy = refs + nh, *bx = refs + 2 * nh, *by = refs + 3 * nh;
    const double *rx = refs + 4 * nh, *ry = refs + 5 * nh, *lx = refs + 6 * nh, *ly = refs + 7 * nh;
    const double* caz = refs + 8 * nh;
    double* outg = P.out + (size_t)b * P.out_stride;

    int status = -1, nactive = 0, bjx1 = 0, bjx2 = 0;
    GiResult res; res.f = 0.0; res.iq = 0; res.status = -1;
    res.it_outer = res.it_add = res.it_drop = res.it_degen = res.it_l2a = 0;
    res.flops = 0;
    bool handed = false;

    bool live = false;
    int i = tick;
    if (!(i < P.gate)) { i -= P.gate; live = (i < P.nsum_mpc - nh); }

    if (!live) {
      for (int k = lane; k < 14; k += 32) outrec[k] = outg[k];
      for (int k = lane; k < 4; k += 32) outrec[14 + k] = theta_in[k];
      for (int k = lane; k < n; k += 32) outrec[18 + k] = xwarm[k];
      if (lane == 0) outrec[18 + n] = 0.0;
    } else {
      bjx1 = duo_indexfind(tx, (i + 1) * dt) + 1;
      bjx2 = duo_indexfind(tx, (i + nh) * dt) + 1;
      const int t_yu = (i + 1) % P.nstepx;
      const bool left = (bjx1 < 2) || (bjx1 % 2 == 0);
      const bool sw = (bjx1 >= 2) && !((t_yu + nh - 1) < P.nstepx);
      const int t_yu_k = (t_yu + nh) - P.nstepx;
      const double thx0 = theta_in[0], thx1 = theta_in[1], thy0 = theta_in[2], thy1 = theta_in[3];
      // ---- cpp:427-526 condensation (lanes own horizon steps; gc0 / s2 are read from the L2-resident model table) ----
      for (int k = lane; k < nh; k += 32) {
        const bool other = sw && (k >= nh - t_yu_k);
        const bool use_l = left ? !other : other;
        const double copx = use_l ? lx[k] : rx[k], copy_ = use_l ? ly[k] : ry[k];
        const double detpx = zx[k] - copx, detpy = zy[k] - copy_;
        const double p = P.j_ini / (P.mass * (caz[k] + P.g));
        pth[k] = p;
        ppsx[k] = fma(pps[k], thx0, pps[nh + k] * thx1);
        ppsy[k] = fma(pps[k], thy0, pps[nh + k] * thy1);
        const double t1x = fma(m1[k], thx0, m1[nh + k] * thx1), t2x = fma(m2[k], thx0, m2[nh + k] * thx1);
        const double t1y = fma(m1[k], thy0, m1[nh + k] * thy1), t2y = fma(m2[k], thy0, m2[nh + k] * thy1);
        double t3x = 0.0, t3y = 0.0;
        for (int j = 0; j < nh; j++) { const double sv = __ldg(g_s2 + j * nh + k); t3x = fma(sv, bx[j], t3x); t3y = fma(sv, by[j], t3y); }
        g0[k] = ((t1x + t2x) - t3x) + (P.gama * p) * detpy;
        g0[nh + k] = ((t1y + t2y) - t3y) + (P.gama * (-p)) * detpx;
      }
      __syncwarp();
      // ---- Hessian block (lower triangle) into R of half 0, trace ----
      double tr = 0.0;
      for (int ii = lane; ii < nh; ii += 32) {
        for (int jj = 0; jj <= ii; jj++) {
          double v = __ldg(g_gc0 + jj * nh + ii);
          if (ii == jj) { v = v + P.gama / 2 * (pth[ii] * pth[ii]); tr += 2 * v; }
          w0.R[jj * ld + ii] = 2 * v;
        }
      }
      tr = warp_sum(tr);
      const double c1 = 2 * tr;
      for (int k = lane; k < n; k += 32) X[k] = xwarm[k];
      for (int t = lane; t < nh * ld; t += 32) w0.J[t] = 0.0;
      __syncwarp();
      res.flops = gi_flops_setup(n, 0);
      if (!gi_llt(w0, nh, lane)) {
        status = ST_NOT_PD;
        res.f = CUDART_INF;
      } else {
        gi_inv_lt(w0, nh, 0, lane);
        double c2 = 0.0;
        for (int k = lane; k < nh; k += 32) c2 += w0.J[k * ld + k];
        c2 = 2 * warp_sum(c2);
        // the two diagonal blocks of G are identical: one factorisation serves both halves
        for (int t = lane; t < nh * ld; t += 32) { w1.J[t] = w0.J[t]; w0.R[t] = 0.0; w1.R[t] = 0.0; }
        // x = -G^-1 g0 = -J (J' g0), half by half
        for (int k = lane; k < nh; k += 32) { w0.np[k] = g0[k]; w1.np[k] = g0[nh + k]; }
        __syncwarp();
        gi_compute_d(w0, 0, nh, lane); gi_update_z(w0, 0, lane);
        gi_compute_d(w1, 0, nh, lane); gi_update_z(w1, 0, lane);
        double f = 0.0;
        for (int k = lane; k < nh; k += 32) {
          const double xa = -w0.z[k], xb = -w1.z[k];
          X[k] = xa; X[nh + k] = xb;
          f = fma(g0[k], xa, f); f = fma(g0[nh + k], xb, f);
        }
        res.f = 0.5 * warp_sum(f);
        __syncwarp();
        DuoShared sh{nh, ppu, ppsx, ppsy, P.j_ini, thmax, P.torque_lim / P.j_ini, X, S};
        const bool ok = duo_loop(w0, w1, sh, XOLD, G, c1, c2, P.cap_scale * (n + m) + 50, res, lane);
        status = res.status;
        nactive = res.iq;
        if (!ok) handed = true;
      }
      if (!handed) {
        // ---- cpp:567-625 first control: fallback / clamp ----
        bool has_nan = false;
        for (int k = lane; k < n; k += 32) has_nan |= (X[k] != X[k]);
        has_nan = __any_sync(FULL_MASK, has_nan);
        if (has_nan && (status == ST_OK || status == ST_EQ_DEP)) status = ST_NAN;
        double ax0 = X[0], ay0 = X[nh];
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
        __syncwarp();
        if (lane == 0) { X[0] = ax0; X[nh] = ay0; }
        __syncwarp();
        // ---- cpp:629-655 roll-out ----
        if (lane < 2) {
          const double* acc = X + lane * nh;
          const double p0 = lane ? thy0 : thx0, v0 = lane ? thy1 : thx1;
          const double a0 = acc[0];
          const double np0 = (p0 + dt * v0) + b0 * a0, nv0 = v0 + b1 * a0;
          const double lam_p = P.lamda[2 * lane], lam_v = P.lamda[2 * lane + 1];
          outrec[14 + 2 * lane] = lam_p * bstate[2 * lane] + (1 - lam_p) * np0;
          outrec[15 + 2 * lane] = lam_v * bstate[2 * lane + 1] + (1 - lam_v) * nv0;
          double pk = p0, vk = v0;
          for (int jj = 0; jj < 3; jj++) {
            const double a = acc[jj];
            const double pn = (pk + dt * vk) + b0 * a, vn = vk + b1 * a;
            pk = pn; vk = vn;
            outrec[(jj == 0 ? 0 : (jj == 1 ? 6 : 10)) + lane] = pk;
          }
          outrec[2 + lane] = P.j_ini * a0;
        }
        if (lane < 3) {
          const int jj = lane;
          const double den = P.mass * (P.g + caz[jj]);
          const double zxr = zx[jj] - P.j_ini * X[nh + jj] / den;
          const double zyr = zy[jj] + P.j_ini * X[jj] / den;
          const int o = (jj == 0) ? 4 : (jj == 1 ? 8 : 12);
          outrec[o] = zxr; outrec[o + 1] = zyr;
        }
        for (int k = lane; k < n; k += 32) outrec[18 + k] = X[k];
        if (lane == 0) outrec[18 + n] = res.f;
      }
    }
    if (handed) {
      // the dense kernel (list mode) takes the instance: nothing of it is written here
      if (lane == 0) {
        const int slot = atomicAdd(P.flist_count, 1);
        if (slot < P.flist_cap) P.flist[slot] = b;
      }
      __syncwarp();
      continue;
    }
    if (lane == 0 && P.out_stride > 19 + n) outrec[19 + n] = 0.0;
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      tma_store_1d(outg, outrec, (uint32_t)(P.out_stride * sizeof(double)));
      tma_store_commit();
    }
    if (P.diag) {
      int* dg = P.diag + (size_t)b * P.diag_stride;
      if (lane == 0) {
        dg[0] = status; dg[1] = nactive;
        dg[2] = res.it_outer; dg[3] = res.it_add; dg[4] = res.it_drop; dg[5] = res.it_degen;
        dg[6] = bjx1; dg[7] = bjx2;
        dg[8] = res.it_l2a; dg[9] = (int)(res.flops > 0x7fffffffull ? 0x7fffffffull : res.flops);
      }
      for (int k = lane; k < n; k += 32) dg[10 + k] = (k < nactive) ? G[k] : -1;
    }
    if (lane == 0) tma_store_wait_read();
    __syncwarp();
  }
  if (lane == 0) tma_store_wait_all();
}

size_t body_duo_smem_bytes(int nh, int wpc, int in_stride, int out_stride, int* warp_doubles) {
  const int wd = duo_warp_doubles(nh, in_stride, out_stride);
  if (warp_doubles) *warp_doubles = wd;
  return (size_t)(duo_cta_doubles(nh) + wpc * wd) * sizeof(double) + (size_t)wpc * sizeof(uint64_t);
}

// picks the warps per CTA that keeps the most warps resident on an SM; grid = persistent over the batch
cudaError_t body_duo_launch(BodyKParams P, int sms, size_t smem_optin, cudaStream_t st) {
  static std::mutex mu;
  static int best_wpc[64][41] = {};
  static int best_occ[64][41] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  int wpc, occ;
  {
    std::lock_guard<std::mutex> lk(mu);
    if (best_wpc[dev][P.nh] == 0) {
      cudaError_t e = cudaFuncSetAttribute(body_duo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_optin);
      if (e != cudaSuccess) return e;
      int bw = 0, bo = 0, bwarps = 0;
      for (int w = 1; w <= DUO_MAX_THREADS / 32; w++) {
        const size_t smem = body_duo_smem_bytes(P.nh, w, P.in_stride, P.out_stride, nullptr);
        if (smem > smem_optin) break;
        int o = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, body_duo_kernel, w * 32, smem);
        if (e != cudaSuccess) return e;
        if (o * w > bwarps) { bwarps = o * w; bw = w; bo = o; }
      }
      if (bw == 0) return cudaErrorLaunchOutOfResources;
      best_wpc[dev][P.nh] = bw; best_occ[dev][P.nh] = bo;
    }
    wpc = best_wpc[dev][P.nh]; occ = best_occ[dev][P.nh];
  }
  int wd = 0;
  const size_t smem = body_duo_smem_bytes(P.nh, wpc, P.in_stride, P.out_stride, &wd);
  P.warp_doubles = wd;
  int grid = (P.B + wpc - 1) / wpc;
  if (grid > sms * occ) grid = sms * occ;
  body_duo_kernel<<<grid, wpc * 32, smem, st>>>(P);
  return cudaGetLastError();
}

}  // namespace go1
