/* TEST INFRASTRUCTURE ONLY (see go1_oracle.h): CPU restatement of the swing-foot generator of the 100 Hz node,
 *   PRMPCClass::Foot_trajectory_solve_mod2   RT/src/FastMPC/PRMPCClass.cpp:1756-2195
 *   PRMPCClass::solve_AAA_inv2               :2224-2236
 *   PRMPCClass::Indexfind (xyz1 = 0 branch)  :716-727
 *   initial members: Initialize :46-55,105-139, FootStepInputs :2198-2221
 * (RT = unitree_ros/rt_mpc_qp), pinned bit for bit against the unmodified class by
 * tests/test_oracle_vs_ref.py::test_oracle_rt_foot_* (live oracle/_ref + tests/golden/rt_foot_ref.npz).
 *
 * gait_fast.cpp:541 calls it every 100 Hz tick once the walk has started, with the step locations / period the 40 Hz
 * planner published (message slots 86-94 = Nrtfoorpr_gen); it returns the right / left foot positions over the body
 * MPC's horizon, of which gait_fast.cpp:585-607 takes x and y as reference rows.
 *
 * Frozen as the reference has them:
 *   - _tx is rebuilt from _ts on every call with the 40 Hz grid rounding (round(tx / 0.025) * 0.025 - 1e-5);
 *     _t_end_footstep uses 2 * tstep here (Initialize uses 3 * tstep);
 *   - _bjxx / _bjx1 keep their last values once j_index > _t_end_footstep;
 *   - only y is written in the `_bjx1 < 2` branch (x, z of that slot keep what an earlier call left there);
 *   - the swing target t_plan(2) = ts - (2 dt_mpc + 0.001) and the landing test |t_des - ts| <= dt_mpc;
 *   - pow() for every monomial, the 4x4 inverse by row-pivoted Gauss-Jordan (what oracle/eigen_shim does above 3x3).
 * Widened: the reference's foot arrays are 1 x 10 (written up to index nh + 1 = 5 at its nh = 4, and read back for 5 steps);
 * here they are nh + 2 wide and nh + 1 steps are returned (= the reference's Vec30 at nh = 4).
 * Clamped: Indexfind's unbounded while loop stops at the end of the 27-entry table.
 */
#include <math.h>
#include <string.h>
#include "go1_oracle.h"

#define NS ORC_FOOTSTEPS

/* state layout (doubles): see go1_oracle.h */
enum { F_TS = 0, F_FX = 27, F_FY = 54, F_FZ = 81, F_LIFT = 108, F_RY = 135, F_BJXX = 136, F_BJX1 = 137, F_ARR = 138 };

void orc_rt_foot_cfg_default(orc_rt_foot_cfg *c)
{
    c->dt = 0.025; c->dt_mpc = 0.01; c->tstep = 0.7; c->tdsp_ratio = 0.1;
    c->stepwidth0 = 0.12675; c->lift_height = 0.015;
}

int orc_rt_foot_state_doubles(int nh) { return F_ARR + 6 * (nh + 2); }

void orc_rt_foot_state_default(const orc_rt_foot_cfg *c, int nh, double *s)
{
    const int W = nh + 2;
    memset(s, 0, sizeof(double) * (size_t)orc_rt_foot_state_doubles(nh));
    /* FootStepInputs(2 * HALF_HIP_WIDTH, 0, 0, lift): _stepwidth(0) is halved, step length / height 0 */
    double sw[NS];
    for (int i = 0; i < NS; i++) { sw[i] = 2 * c->stepwidth0; s[F_TS + i] = c->tstep; s[F_LIFT + i] = c->lift_height; }
    sw[0] = sw[0] / 2;
    s[F_LIFT + NS - 1] = 0; s[F_LIFT + NS - 2] = 0; s[F_LIFT + NS - 3] = c->lift_height / 2; s[F_LIFT + NS - 4] = c->lift_height;
    for (int i = 1; i < NS; i++) s[F_FY + i] = s[F_FY + i - 1] + (int)pow(-1, i - 1) * sw[i - 1];
    double *Ry = s + F_ARR + 1 * W, *Ly = s + F_ARR + 4 * W;
    for (int k = 0; k < W; k++) { Ry[k] = -sw[0]; Ly[k] = sw[0]; }
}

static void gj_inverse4(double *a, double *r)
{
    const int n = 4;
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

/* out: 6 * (nh + 1) doubles, per step (R x y z, L x y z). */
void orc_rt_foot_traj(const orc_rt_foot_cfg *c, int nh, double *s, int j_indexx, int stopwalking, const double nrt[9], double *out)
{
    const int W = nh + 2;
    const double dt_mpc = c->dt_mpc;
    double *ts = s + F_TS, *fxyz[3] = { s + F_FX, s + F_FY, s + F_FZ }, *lift = s + F_LIFT;
    double *Rf[3] = { s + F_ARR, s + F_ARR + W, s + F_ARR + 2 * W }, *Lf[3] = { s + F_ARR + 3 * W, s + F_ARR + 4 * W, s + F_ARR + 5 * W };
    /* :1758-1770 step locations / period published by the 40 Hz planner */
    const int bjxx_nrt = (int)nrt[0];
    fxyz[0][bjxx_nrt] = nrt[1]; fxyz[0][bjxx_nrt + 1] = nrt[2];
    fxyz[1][bjxx_nrt] = nrt[3]; fxyz[1][bjxx_nrt + 1] = nrt[4];
    fxyz[2][bjxx_nrt] = nrt[5]; fxyz[2][bjxx_nrt + 1] = nrt[6];
    const int bjx_period_nrt = (int)nrt[7];
    if (nrt[8] > 0) ts[bjx_period_nrt] = nrt[8];
    /* :1774-1784 */
    double td[NS], tx[NS];
    for (int i = 0; i < NS; i++) td[i] = c->tdsp_ratio * ts[i];
    tx[0] = 0;
    for (int i = 1; i < NS; i++) { tx[i] = tx[i - 1] + ts[i - 1]; tx[i] = round(tx[i] / c->dt) * c->dt - 0.00001; }
    const double t_end = round((tx[NS - 1] - 2 * c->tstep) / dt_mpc);
    int bjxx = (int)s[F_BJXX], bjx1 = (int)s[F_BJX1];
    double ry = s[F_RY];

    for (int j_index = j_indexx; j_index < j_indexx + nh; j_index++) {
        const int q = j_index - j_indexx;          /* slot of the sample before: the arrays are written at q + 1 and q + 2 */
        if (j_index <= t_end) {
            int jp = 0;
            while (jp < NS && j_index * dt_mpc >= tx[jp]) jp++;
            bjxx = (jp - 1) + 1;
            jp = 0;
            while (jp < NS && (j_index + 1) * dt_mpc >= tx[jp]) jp++;
            bjx1 = (jp - 1) + 1;
        }
        if (stopwalking || (j_index > t_end))
            for (int it = bjx1 + 1; it < NS; it++) lift[it] = 0;
        for (int it = 24; it < NS; it++) lift[it] = 0;
        fxyz[1][0] = -c->stepwidth0;

        if ((bjx1 >= 2) && (j_index <= t_end)) {
            double **sup = (bjx1 % 2 == 0) ? Lf : Rf;     /* even: left support, right swing */
            double **swg = (bjx1 % 2 == 0) ? Rf : Lf;
            for (int k = 0; k < 3; k++) { sup[k][q + 1] = sup[k][q]; sup[k][q + 2] = sup[k][q]; }
            const double s0 = round(tx[bjx1 - 1] / dt_mpc);
            if ((j_index + 1 - s0) * dt_mpc < td[bjx1 - 1]) {
                for (int k = 0; k < 3; k++) { swg[k][q + 1] = swg[k][q]; swg[k][q + 2] = swg[k][q]; }
            } else {
                const double t_des = (j_index + 1 - s0 + 1) * dt_mpc;
                const double tp[3] = { t_des - dt_mpc, (td[bjx1 - 1] + ts[bjx1 - 1]) / 2 + 0.0001, ts[bjx1 - 1] - (2 * dt_mpc + 0.001) };
                if (fabs(t_des - ts[bjx1 - 1]) <= (dt_mpc)) {
                    for (int k = 0; k < 3; k++) { swg[k][q + 1] = fxyz[k][bjxx]; swg[k][q + 2] = fxyz[k][bjxx]; }
                } else {
                    double A[16], Ai[16];
                    for (int r = 0; r < 3; r++) { A[4 * r] = pow(tp[r], 3); A[4 * r + 1] = pow(tp[r], 2); A[4 * r + 2] = pow(tp[r], 1); A[4 * r + 3] = 1; }
                    A[12] = 3 * pow(tp[2], 2); A[13] = 2 * pow(tp[2], 1); A[14] = pow(tp[2], 0); A[15] = 0;
                    gj_inverse4(A, Ai);
                    const double tap[4] = { pow(t_des, 3), pow(t_des, 2), pow(t_des, 1), 1 };
                    const double tav[4] = { 3 * pow(t_des, 2), 2 * pow(t_des, 1), 1, 0 };
                    if ((j_index + 1 - s0) * dt_mpc < td[bjx1 - 1] + dt_mpc)
                        ry = (fxyz[1][bjxx] + fxyz[1][bjxx - 2]) / 2;
                    for (int k = 0; k < 3; k++) {
                        double plan[4];
                        plan[0] = swg[k][q];
                        if (k == 0) plan[1] = (fxyz[0][bjxx - 2] + fxyz[0][bjxx]) / 2;
                        else if (k == 1) plan[1] = ry;
                        else plan[1] = fmax(fxyz[2][bjxx - 2], fxyz[2][bjxx]) + lift[bjx1 - 1];
                        plan[2] = fxyz[k][bjxx];
                        plan[3] = 0;
                        double co[4];
                        for (int r = 0; r < 4; r++) { double a_ = 0.0; for (int m = 0; m < 4; m++) a_ += Ai[4 * r + m] * plan[m]; co[r] = a_; }
                        double p_ = 0.0, v_ = 0.0;
                        for (int m = 0; m < 4; m++) { p_ += tap[m] * co[m]; v_ += tav[m] * co[m]; }
                        swg[k][q + 1] = p_;
                        swg[k][q + 2] = p_ + dt_mpc * v_;
                    }
                }
            }
        } else {
            if (j_index > t_end) {
                for (int k = 0; k < 3; k++) { Rf[k][q + 1] = Rf[k][q]; Lf[k][q + 1] = Lf[k][q]; }
            } else {
                Rf[1][q + 1] = -c->stepwidth0;
                Lf[1][q + 1] = c->stepwidth0;
            }
        }
    }
    for (int jjj = 0; jjj < nh + 1; jjj++)
        for (int k = 0; k < 3; k++) { out[6 * jjj + k] = Rf[k][jjj + 1]; out[6 * jjj + 3 + k] = Lf[k][jjj + 1]; }
    for (int k = 0; k < 3; k++) { Rf[k][0] = Rf[k][1]; Lf[k][0] = Lf[k][1]; }
    s[F_BJXX] = bjxx; s[F_BJX1] = bjx1; s[F_RY] = ry;
}
