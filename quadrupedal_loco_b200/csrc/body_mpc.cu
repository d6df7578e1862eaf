// body_mpc.cu -- fused body-inclination MPC tick: condensation -> QP -> clamp -> roll-out.
//
// Replaces PRMPCClass::body_theta_mpc (RT/src/FastMPC/PRMPCClass.cpp:379-714) with
// solve_body_rotation/Solve (:799-849) and Indexfind (:716-738) for a batch of
// independent instances, horizon nh (reference: compile-time 4).
//
// One warp per instance.  The instance's input record (tx | tick | theta |
// bodyangle_state | x_warm | 9 reference rows) arrives in shared memory by one TMA
// bulk copy; the horizon model (Ppu, the constant part of the Hessian, beta*Ppu',
// the state-propagation products) is shared by the CTA and loaded once by a second
// bulk copy.  The dense QP data the reference builds per tick -- G (2nh x 2nh), CI
// (2nh x 12nh), ci0 -- is never formed:
//   * G is block diagonal with two IDENTICAL nh x nh blocks (pthetay = -pthetax,
//     equal weights), so one nh x nh Cholesky gives L and J = L^-T for both;
//   * CI's columns are +-rows of [Ppu 0], [0 Ppu] and +-j_ini e_k; n+ and
//     s = CI'x + ci0 are generated from Ppu on the fly (BodyPolicy), and the 4nh
//     never-populated columns (cpp:813-816) only enter through m in the solver's
//     stop tolerance.
// HBM traffic per instance: in_stride + 2*out_stride doubles + diag ints.
#include <cuda_runtime.h>
#include <stdint.h>
#include "gi_warp.cuh"
#include "tma.cuh"
#include "kernels.h"

namespace go1 {

struct BodyPolicy {
  int nh;
  const double* ppu;   // shared, nh x nh column-major (lower triangular)
  const double* ppsx;  // shared, nh: Pps * thetaxk
  const double* ppsy;
  double j_ini, thmax, tq;

  // s for constraint c < 8nh given x
  __device__ __forceinline__ double s_of(const GiWs& w, int c) const {
    int blk = c / nh, k = c - blk * nh;
    if (blk < 4) {
      int half = blk >> 1, low = blk & 1;
      const double* xx = w.x + half * nh;
      double v = 0.0;
      for (int j = 0; j <= k; j++) v = fma(ppu[j * nh + k], xx[j], v);
      double pk = (half ? ppsy : ppsx)[k];
      return low ? (v + (thmax + pk)) : ((thmax - pk) - v);
    }
    int b2 = blk - 4, half = b2 >> 1, low = b2 & 1;
    double xv = w.x[half * nh + k];
    return low ? fma(j_ini, xv, tq) : fma(-j_ini, xv, tq);
  }
  __device__ __forceinline__ void eval_s(const GiWs& w, int lane, double& psi) const {
    for (int c = lane; c < 8 * nh; c += 32) {
      double sv = s_of(w, c);
      w.s[c] = sv;
      psi += fmin(0.0, sv);
    }
  }
  __device__ __forceinline__ void load_np(const GiWs& w, int ip, int lane, int& klo, int& khi) const {
    int blk = ip / nh, k = ip - blk * nh;
    int half, low, lo, hi;
    if (blk < 4) { half = blk >> 1; low = blk & 1; lo = half * nh; hi = lo + k + 1; }
    else { int b2 = blk - 4; half = b2 >> 1; low = b2 & 1; lo = half * nh + k; hi = lo + 1; }
    for (int j = lane; j < w.n; j += 32) {
      double v = 0.0;
      if (j >= lo && j < hi) {
        v = (blk < 4) ? ppu[(j - half * nh) * nh + k] : j_ini;
        if (!low) v = -v;
      }
      w.np[j] = v;
    }
    klo = lo; khi = hi;
    __syncwarp();
  }
  __device__ __forceinline__ double eval_one(const GiWs& w, int ip, int) const { return s_of(w, ip); }
  __device__ __forceinline__ void load_eq(const GiWs&, int, int, bool& allzero) const { allzero = true; }
  __device__ __forceinline__ double ce0(int) const { return 0.0; }
};

// PRMPCClass::Indexfind, xyz < 0.05 branch (cpp:716-727), clamped to the 27-entry table
__device__ __forceinline__ int body_indexfind(const double* tx, double goal) {
  int j = 0;
  while (j < 27 && goal >= tx[j]) j++;
  return j - 1;
}

template <int WPC>
__global__ void __launch_bounds__(WPC * 32) body_mpc_kernel(BodyKParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* smem = reinterpret_cast<double*>(smem_raw);
  const int nh = P.nh, n = 2 * nh, m = 12 * nh;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // CTA-shared model table
  double* tab = smem;
  const double* ppu = tab;                   // nh x nh
  const double* gc0 = tab + nh * nh;         // nh x nh   (R/2 I + alpha/2 Pvu'Pvu) + beta/2 Ppu'Ppu
  const double* s2 = tab + 2 * nh * nh;      // nh x nh   beta * Ppu'
  const double* m1 = tab + 3 * nh * nh;      // nh x 2    (alpha Pvu') Pvs
  const double* m2 = m1 + 2 * nh;            // nh x 2    (beta Ppu') Pps
  const double* pps = m2 + 2 * nh;           // nh x 2
  // per-warp slice
  double* wbase = smem + P.tab_doubles + (size_t)warp * P.warp_doubles;
  GiWs w;
  gi_ws_carve(w, wbase, n, 0, 8 * nh);       // s[] only for the 8nh populated rows
  w.m = m; w.ms = 8 * nh;
  double* inrec = wbase + gi_ws_doubles(n, 8 * nh);
  double* outrec = inrec + P.in_stride;
  double* pth = outrec + P.out_stride;       // nh
  double* ppsx = pth + nh;                   // nh
  double* ppsy = ppsx + nh;                  // nh
  double* g0 = ppsy + nh;                    // 2nh
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P.tab_doubles + (size_t)WPC * P.warp_doubles);
  uint64_t* tab_bar = bars + WPC;
  uint64_t* my_bar = bars + warp;

  if (threadIdx.x == 0) {
    for (int i = 0; i <= WPC; i++) mbar_init(bars + i, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(tab_bar, (uint32_t)(P.tab_doubles * sizeof(double)));
    tma_load_1d(tab, P.tab, (uint32_t)(P.tab_doubles * sizeof(double)), tab_bar);
  }
  bool tab_ready = false;
  uint32_t phase = 0;

  const double dt = P.dt_mpc;
  const double b0 = dt * dt / 2, b1 = dt;  // _b = [dt^2/2, dt]; pow(dt,2) == dt*dt exactly
  const double thmax = P.theta_lim, thmin = -P.theta_lim;

  // list mode (P.flist != null): the instances body_duo.cu handed over -- or all B when the list overflowed
  int total = P.B;
  bool listed = false;
  if (P.flist) {
    const int cnt = *P.flist_count;
    listed = cnt <= P.flist_cap;
    total = listed ? cnt : P.B;
    if (blockIdx.x == 0 && threadIdx.x == 0 && cnt > 0) atomicAdd(P.flist_count + 1, cnt);    // hand-over statistic
  }
  for (int bi = blockIdx.x * WPC + warp; bi < total; bi += gridDim.x * WPC) {
    const int b = listed ? P.flist[bi] : bi;
    // ---- stage the input record (TMA bulk copy, async proxy) ----
    if (lane == 0) {
      mbar_expect_tx(my_bar, (uint32_t)(P.in_stride * sizeof(double)));
      tma_load_1d(inrec, P.in + (size_t)b * P.in_stride, (uint32_t)(P.in_stride * sizeof(double)), my_bar);
    }
    __syncwarp();
    if (!tab_ready) { mbar_wait(tab_bar, 0); tab_ready = true; }
    mbar_wait(my_bar, phase);
    phase ^= 1u;

    const double* tx = inrec;
    int tick = (int)inrec[27];
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

    bool live = false;
    int i = tick;
    if (!(i < P.gate)) { i -= P.gate; live = (i < P.nsum_mpc - nh); }

    if (!live) {
      // gated tick: the reference returns its stale members; state and V_ini unchanged
      for (int k = lane; k < 14; k += 32) outrec[k] = outg[k];
      for (int k = lane; k < 4; k += 32) outrec[14 + k] = theta_in[k];
      for (int k = lane; k < n; k += 32) outrec[18 + k] = xwarm[k];
      if (lane == 0) outrec[18 + n] = 0.0;
    } else {
      // ---- cpp:406-417 phase indices ----
      bjx1 = body_indexfind(tx, (i + 1) * dt) + 1;
      bjx2 = body_indexfind(tx, (i + nh) * dt) + 1;
      const int t_yu = (i + 1) % P.nstepx;
      const bool left = (bjx1 < 2) || (bjx1 % 2 == 0);
      const bool sw = (bjx1 >= 2) && !((t_yu + nh - 1) < P.nstepx);
      const int t_yu_k = (t_yu + nh) - P.nstepx;
      const double thx0 = theta_in[0], thx1 = theta_in[1], thy0 = theta_in[2], thy1 = theta_in[3];

      // ---- cpp:427-526 condensation (lanes own horizon steps) ----
      for (int k = lane; k < nh; k += 32) {
        bool other = sw && (k >= nh - t_yu_k);   // the tail of the window is on the other foot
        bool use_l = left ? !other : other;
        double copx = use_l ? lx[k] : rx[k], copy_ = use_l ? ly[k] : ry[k];
        double detpx = zx[k] - copx, detpy = zy[k] - copy_;
        double p = P.j_ini / (P.mass * (caz[k] + P.g));
        pth[k] = p;
        double px = fma(pps[k], thx0, pps[nh + k] * thx1);
        double py = fma(pps[k], thy0, pps[nh + k] * thy1);
        ppsx[k] = px; ppsy[k] = py;
        // gradient: ((alpha Pvu' Pvs + beta Ppu' Pps) theta - beta Ppu' ref) + gama ptheta det
        double t1x = fma(m1[k], thx0, m1[nh + k] * thx1), t2x = fma(m2[k], thx0, m2[nh + k] * thx1);
        double t1y = fma(m1[k], thy0, m1[nh + k] * thy1), t2y = fma(m2[k], thy0, m2[nh + k] * thy1);
        double t3x = 0.0, t3y = 0.0;
        for (int j = 0; j < nh; j++) { t3x = fma(s2[j * nh + k], bx[j], t3x); t3y = fma(s2[j * nh + k], by[j], t3y); }
        g0[k] = ((t1x + t2x) - t3x) + (P.gama * p) * detpy;
        g0[nh + k] = ((t1y + t2y) - t3y) + (P.gama * (-p)) * detpx;
      }
      __syncwarp();
      // ---- Hessian block (lower triangle) into R, trace ----
      double tr = 0.0;
      for (int ii = lane; ii < nh; ii += 32) {
        for (int jj = 0; jj <= ii; jj++) {
          double v = gc0[jj * nh + ii];
          if (ii == jj) { v = v + P.gama / 2 * (pth[ii] * pth[ii]); tr += 2 * v; }
          w.R[jj * w.ld + ii] = 2 * v;
        }
      }
      tr = warp_sum(tr);
      const double c1 = 2 * tr;
      __syncwarp();
      for (int k = lane; k < n; k += 32) w.x[k] = xwarm[k];
      for (int t = lane; t < n * w.ld; t += 32) w.J[t] = 0.0;
      __syncwarp();

      res.flops = gi_flops_setup(n, 0);
      if (!gi_llt(w, nh, lane)) {
        status = ST_NOT_PD;   // x keeps the warm start, exactly as the reference
        res.f = CUDART_INF;
      } else {
        gi_inv_lt(w, nh, 0, lane);
        gi_inv_lt(w, nh, nh, lane);
        double c2 = 0.0;
        for (int k = lane; k < nh; k += 32) c2 += w.J[k * w.ld + k];
        c2 = 2 * warp_sum(c2);
        for (int t = lane; t < n * w.ld; t += 32) w.R[t] = 0.0;
        // x = -G^-1 g0 = -J (J' g0)
        for (int k = lane; k < n; k += 32) w.np[k] = g0[k];
        __syncwarp();
        gi_compute_d(w, 0, n, lane);
        gi_update_z(w, 0, lane);
        double f = 0.0;
        for (int k = lane; k < n; k += 32) { double xv = -w.z[k]; w.x[k] = xv; f = fma(g0[k], xv, f); }
        res.f = 0.5 * warp_sum(f);
        __syncwarp();
        BodyPolicy pol{nh, ppu, ppsx, ppsy, P.j_ini, thmax, P.torque_lim / P.j_ini};
        gi_loop(w, pol, c1, c2, P.cap_scale * (n + m) + 50, res, lane);
        status = res.status;
        nactive = res.iq;
      }
      // ---- cpp:567-625 first control: fallback / clamp ----
      bool has_nan = false;
      for (int k = lane; k < n; k += 32) has_nan |= (w.x[k] != w.x[k]);
      has_nan = __any_sync(FULL_MASK, has_nan);
      if (has_nan && (status == ST_OK || status == ST_EQ_DEP)) status = ST_NAN;
      double ax0 = w.x[0], ay0 = w.x[nh];
      const double arow_x = thx0 + dt * thx1, arow_y = thy0 + dt * thy1;   // _a.row(0) * theta
      if (has_nan) {
        ax0 = (thx0 - arow_x) / b0;
        ay0 = (thy0 - arow_y) / b0;
      } else {
        double nx0 = arow_x + b0 * ax0;
        if (nx0 > thmax) ax0 = (thmax - arow_x) / b0;
        else if (nx0 < thmin) ax0 = (thmin - arow_x) / b0;
        double ny0 = arow_y + b0 * ay0;
        if (ny0 > thmax) ay0 = (thmax - arow_y) / b0;
        else if (ny0 < thmin) ay0 = (thmin - arow_y) / b0;
      }
      __syncwarp();
      if (lane == 0) { w.x[0] = ax0; w.x[nh] = ay0; }
      __syncwarp();
      // ---- cpp:629-655 roll-out: lane 0 integrates roll, lane 1 pitch; out14 needs steps 0..2 ----
      if (lane < 2) {
        const double* acc = w.x + lane * nh;
        double p0 = lane ? thy0 : thx0, v0 = lane ? thy1 : thx1;
        double a0 = acc[0];
        double np0 = (p0 + dt * v0) + b0 * a0, nv0 = v0 + b1 * a0;   // state after the first control
        double lam_p = P.lamda[2 * lane], lam_v = P.lamda[2 * lane + 1];
        outrec[14 + 2 * lane] = lam_p * bstate[2 * lane] + (1 - lam_p) * np0;
        outrec[15 + 2 * lane] = lam_v * bstate[2 * lane + 1] + (1 - lam_v) * nv0;
        double pk = p0, vk = v0;
        for (int jj = 0; jj < 3; jj++) {
          double a = acc[jj];
          double pn = (pk + dt * vk) + b0 * a, vn = vk + b1 * a;
          pk = pn; vk = vn;
          outrec[(jj == 0 ? 0 : (jj == 1 ? 6 : 10)) + lane] = pk;   // thetax/thetay at k+1, k+2, k+3
        }
        outrec[2 + lane] = P.j_ini * a0;                            // torque
      }
      if (lane < 3) {
        // cpp:651-652 ZMP consistent with the planned angular acceleration
        int jj = lane;
        double den = P.mass * (P.g + caz[jj]);
        double zxr = zx[jj] - P.j_ini * w.x[nh + jj] / den;
        double zyr = zy[jj] + P.j_ini * w.x[jj] / den;
        int o = (jj == 0) ? 4 : (jj == 1 ? 8 : 12);
        outrec[o] = zxr; outrec[o + 1] = zyr;
      }
      for (int k = lane; k < n; k += 32) outrec[18 + k] = w.x[k];
      if (lane == 0) outrec[18 + n] = res.f;
    }
    if (lane == 0 && P.out_stride > 19 + n) outrec[19 + n] = 0.0;   // pad double of the record
    // ---- write back: one TMA bulk store of the output record, diag by lanes ----
    fence_proxy_async();   // every lane: its generic-proxy writes to outrec become visible to the async proxy
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
      for (int k = lane; k < n; k += 32) dg[10 + k] = (k < nactive) ? w.A[k] : -1;
    }
    if (lane == 0) tma_store_wait_read();   // outrec may be overwritten by the next instance
    __syncwarp();
  }
  if (lane == 0) tma_store_wait_all();
}

size_t body_smem_bytes(int nh, int wpc, int in_stride, int out_stride, int tab_doubles, int* warp_doubles) {
  int n = 2 * nh;
  int wd = gi_ws_doubles(n, 8 * nh) + in_stride + out_stride + 5 * nh;
  wd = (wd + 1) & ~1;
  if (warp_doubles) *warp_doubles = wd;
  return (size_t)(tab_doubles + wpc * wd) * sizeof(double) + (size_t)(wpc + 1) * sizeof(uint64_t);
}

template <int WPC>
static cudaError_t launch_wpc(const BodyKParams& P, int grid, size_t smem, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(body_mpc_kernel<WPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  body_mpc_kernel<WPC><<<grid, WPC * 32, smem, st>>>(P);
  return cudaGetLastError();
}

cudaError_t body_mpc_launch(BodyKParams P, int wpc, int grid, size_t smem, cudaStream_t st) {
  switch (wpc) {
    case 1: return launch_wpc<1>(P, grid, smem, st);
    case 2: return launch_wpc<2>(P, grid, smem, st);
    case 4: return launch_wpc<4>(P, grid, smem, st);
    case 8: return launch_wpc<8>(P, grid, smem, st);
    default: return cudaErrorInvalidValue;
  }
}

template <int WPC>
static cudaError_t occ_wpc(int* blocks, size_t smem) {
  cudaError_t e = cudaFuncSetAttribute(body_mpc_kernel<WPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, body_mpc_kernel<WPC>, WPC * 32, smem);
}
cudaError_t body_mpc_occupancy(int wpc, size_t smem, int* blocks_per_sm) {
  switch (wpc) {
    case 1: return occ_wpc<1>(blocks_per_sm, smem);
    case 2: return occ_wpc<2>(blocks_per_sm, smem);
    case 4: return occ_wpc<4>(blocks_per_sm, smem);
    case 8: return occ_wpc<8>(blocks_per_sm, smem);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace go1
