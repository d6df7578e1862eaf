/*
 * body_mpc.c -- oracle restatement of the body-inclination (roll/pitch) MPC.
 *
 * TEST INFRASTRUCTURE (see go1_oracle.h).  Restates, with the horizon as a
 * run-time parameter (the reference fixes _nh = 4, PRMPCClass.h:34):
 *   PRMPCClass::Initialize (model part)  RT/src/FastMPC/PRMPCClass.cpp:166-261,280-287,356-360
 *   PRMPCClass::Matrix_ps / Matrix_pu    RT/src/FastMPC/PRMPCClass.cpp:741-796
 *   PRMPCClass::body_theta_mpc           RT/src/FastMPC/PRMPCClass.cpp:379-714
 *   PRMPCClass::Indexfind                RT/src/FastMPC/PRMPCClass.cpp:716-738
 *   PRMPCClass::solve_body_rotation/Solve RT/src/FastMPC/PRMPCClass.cpp:799-849
 * The Eigen expressions are evaluated in the order the expression tree gives
 * (left-associated products, element-wise sums), each product accumulating in
 * ascending inner index from 0 -- the order oracle/eigen_shim uses, so that this
 * file and the unmodified reference compiled against the shim agree bit for
 * bit at nh = 4 (tests/test_oracle_vs_ref.py).
 *
 * Frozen quirks: CI/ci0 columns 8nh..12nh-1 are never written by the reference
 * (cpp:813-816,826-829) and are zero here, but m stays 12nh because it enters
 * the solver's stop tolerance; the warm start X = V_ini is a no-op unless G is
 * not PD; gated ticks return the stale rolled-out members; the QP "success"
 * flag is only "no NaN in X" (QPBaseClass.cpp:137-148), so an infeasible
 * solve's x is consumed.
 * Differences: Indexfind's unbounded while (cpp:721-724) is clamped to the 27
 * table entries; cout prints are dropped.
 */
#include <math.h>
#include <string.h>
#include "go1_oracle.h"

/* C = A(m x k) * B(k x n), column-major, acc from 0 in ascending k */
static void mm(const double *A, const double *B, double *C, int m, int k, int n)
{
    for (int j = 0; j < n; j++)
        for (int i = 0; i < m; i++) {
            double acc = 0.0;
            for (int t = 0; t < k; t++) acc += A[t * m + i] * B[j * k + t];
            C[j * m + i] = acc;
        }
}

void orc_body_cfg_default(orc_body_cfg *c, int nh)
{
    /* RT/src/Robotpara/robot_const_para_config.cpp:8-47, PRMPCClass.cpp:28-42,157-170 */
    c->nh = nh;
    c->dt_mpc = 0.01;
    c->dt_slow = 0.025;
    c->tstep = 0.7;
    c->height_offset_time = 1.0;
    c->g = 9.8;
    c->mass = 12.0;
    c->j_ini = 12 * 0.1 * 0.1;
    c->foot_length = 0.02;
    c->foot_width = 0.02;
    c->theta_lim = 10 * M_PI / 180;
    c->torque_lim = 20.0;
    c->Rtheta = 100.0;
    c->alphatheta = 10.0;
    c->beltatheta = 5000000000.0;
    c->gama_zmp = 5000.0;
    c->lamda[0] = c->lamda[1] = c->lamda[2] = c->lamda[3] = 0.0;
}

/* Matrix_ps: row i = cx * a^(i+1)   (cpp:741-763) */
static void matrix_ps(const double a[4], int nh, const double cx[2], double *out /* nh x 2 */)
{
    for (int i = 0; i < nh; i++) {
        double A[4] = { 1, 0, 0, 1 }, T[4];
        for (int j = 1; j < i + 2; j++) { mm(A, a, T, 2, 2, 2); memcpy(A, T, sizeof A); }
        double r[2];
        mm(cx, A, r, 1, 2, 2);
        out[0 * nh + i] = r[0];
        out[1 * nh + i] = r[1];
    }
}

/* Matrix_pu: (i,j) = cx * a^(i-j) * b, lower triangular   (cpp:765-796) */
static void matrix_pu(const double a[4], const double b[2], int nh, const double cx[2], double *out)
{
    memset(out, 0, sizeof(double) * nh * nh);
    for (int i = 1; i < nh + 1; i++)
        for (int j = 1; j < i + 1; j++) {
            double A[4] = { 1, 0, 0, 1 }, T[4];
            if (j != i)
                for (int k = 1; k < i - j + 1; k++) { mm(A, a, T, 2, 2, 2); memcpy(A, T, sizeof A); }
            double r[2], v[1];
            mm(cx, A, r, 1, 2, 2);
            mm(r, b, v, 1, 2, 1);
            out[(j - 1) * nh + (i - 1)] = v[0];
        }
}

void orc_body_init(orc_body_mpc *s, const orc_body_cfg *c)
{
    memset(s, 0, sizeof *s);
    s->cfg = *c;
    int nh = c->nh;
    /* cpp:168-185 step tables */
    s->nstepx = (int)round(c->tstep / c->dt_mpc);
    s->tx[0] = 0.0;
    for (int i = 1; i < ORC_FOOTSTEPS; i++) {
        s->tx[i] = s->tx[i - 1] + c->tstep;
        s->tx[i] = round(s->tx[i] / c->dt_slow) * c->dt_slow - 0.00001;
    }
    s->nsum_mpc = (int)floor(s->tx[ORC_FOOTSTEPS - 1] / c->dt_mpc);
    /* cpp:198-220 double-integrator prediction model */
    double a[4] = { 1, 0, c->dt_mpc, 1 };               /* column-major [[1,dt],[0,1]] */
    double b[2] = { pow(c->dt_mpc, 2) / 2, c->dt_mpc };
    double cp[2] = { 1, 0 }, cv[2] = { 0, 1 };
    matrix_ps(a, nh, cp, s->pps);
    matrix_ps(a, nh, cv, s->pvs);
    matrix_pu(a, b, nh, cp, s->ppu);
    matrix_pu(a, b, nh, cv, s->pvu);
    double T[ORC_BODY_NH_MAX * ORC_BODY_NH_MAX];
    for (int j = 0; j < nh; j++) for (int i = 0; i < nh; i++) T[j * nh + i] = s->pvu[i * nh + j];
    mm(T, s->pvu, s->pvu_2, nh, nh, nh);
    for (int j = 0; j < nh; j++) for (int i = 0; i < nh; i++) T[j * nh + i] = s->ppu[i * nh + j];
    mm(T, s->ppu, s->ppu_2, nh, nh, nh);
    s->qp_solution = 1;
}

/* cpp:716-727 (xyz < 0.05 branch), clamped to the table */
static int body_indexfind(const double *tx, double goal)
{
    int j = 0;
    while (j < ORC_FOOTSTEPS && goal >= tx[j]) j++;
    return j - 1;
}

void orc_body_theta_mpc(orc_body_mpc *s, int i, const double bodyangle_state[4],
                        const double *zmp_ref, const double *bodyangle_ref,
                        const double *rfoot_ref, const double *lfoot_ref,
                        const double *comacc_z_ref, double out14[14])
{
    const orc_body_cfg *c = &s->cfg;
    const int nh = c->nh, n = 2 * nh, m = 12 * nh;
    const double a[4] = { 1, 0, c->dt_mpc, 1 };
    const double b[2] = { pow(c->dt_mpc, 2) / 2, c->dt_mpc };
    const double thmax = c->theta_lim, thmin = -c->theta_lim;
    const double tqmax = c->torque_lim / c->j_ini, tqmin = -c->torque_lim / c->j_ini;
    const double zmpx_max = c->foot_length / 2 + 0, zmpx_min = -(c->foot_length / 2 - 0);
    const double zmpy_max = c->foot_width / 2, zmpy_min = -c->foot_width / 2;
    const double *zx = zmp_ref, *zy = zmp_ref + nh;
    const double *bx = bodyangle_ref, *by = bodyangle_ref + nh;
    const double *rx = rfoot_ref, *ry = rfoot_ref + nh;
    const double *lx = lfoot_ref, *ly = lfoot_ref + nh;

    int gate = (int)round(c->height_offset_time / c->dt_mpc);
    if (!(i < gate)) {
        i -= gate;
        if (i < s->nsum_mpc - nh) {
            /* cpp:406-417 */
            double tf0 = (i + 1) * c->dt_mpc, tfN = (i + nh) * c->dt_mpc;
            s->bjx1 = body_indexfind(s->tx, tf0) + 1;
            s->bjx2 = body_indexfind(s->tx, tfN) + 1;
            int t_yu = (i + 1) % s->nstepx;

            double zxmax[ORC_BODY_NH_MAX], zxmin[ORC_BODY_NH_MAX], zymax[ORC_BODY_NH_MAX], zymin[ORC_BODY_NH_MAX];
            double copx[ORC_BODY_NH_MAX], copy_[ORC_BODY_NH_MAX];
            for (int k = 0; k < nh; k++) { zxmax[k] = zmpx_max; zxmin[k] = zmpx_min; zymax[k] = zmpy_max; zymin[k] = zmpy_min; }
            /* cpp:427-499 CoP centre reference */
            const double *sx, *sy, *ox, *oy;   /* support foot, other foot */
            int left = (s->bjx1 < 2) || (s->bjx1 % 2 == 0);
            if (left) { sx = lx; sy = ly; ox = rx; oy = ry; } else { sx = rx; sy = ry; ox = lx; oy = ly; }
            for (int k = 0; k < nh; k++) { copx[k] = sx[k]; copy_[k] = sy[k]; }
            if (s->bjx1 >= 2 && !((t_yu + nh - 1) < s->nstepx)) {
                int t_yu_k = (t_yu + nh) - s->nstepx;
                for (int jx = 1; jx <= t_yu_k; jx++) {
                    int k = nh - jx;
                    copx[k] = ox[k];
                    copy_[k] = oy[k];
                    zxmax[k] = fmax(rx[k], lx[k]) - ox[k] + zmpx_max;
                    zxmin[k] = fmin(rx[k], lx[k]) - ox[k] + zmpx_min;
                    zymax[k] = fmax(ry[k], ly[k]) - oy[k] + zmpy_max;
                    zymin[k] = fmin(ry[k], ly[k]) - oy[k] + zmpy_min;
                }
            }
            (void)zxmax; (void)zxmin; (void)zymax; (void)zymin; /* ZMP rows are built but not inserted (cpp:813-816) */

            /* cpp:504-515 condensation: diagonal inertia term + constant Gram matrices */
            double pth[ORC_BODY_NH_MAX];
            for (int k = 0; k < nh; k++) pth[k] = c->j_ini / (c->mass * (comacc_z_ref[k] + c->g));
            double G[4 * ORC_BODY_NH_MAX * ORC_BODY_NH_MAX];
            memset(G, 0, sizeof(double) * n * n);
            for (int jj = 0; jj < nh; jj++)
                for (int ii = 0; ii < nh; ii++) {
                    double unit = (ii == jj) ? 1.0 : 0.0;
                    /* (pthetax * pthetax') and (pthetay * pthetay') with pthetay = -pthetax:
                     * full products over a diagonal matrix reduce exactly to p_i*p_i on the diagonal */
                    double ppx = (ii == jj) ? pth[ii] * pth[ii] : 0.0;
                    double ppy = (ii == jj) ? (-pth[ii]) * (-pth[ii]) : 0.0;
                    double wx = c->Rtheta / 2 * unit + c->alphatheta / 2 * s->pvu_2[jj * nh + ii]
                              + c->beltatheta / 2 * s->ppu_2[jj * nh + ii] + c->gama_zmp / 2 * ppx;
                    double wy = c->Rtheta / 2 * unit + c->alphatheta / 2 * s->pvu_2[jj * nh + ii]
                              + c->beltatheta / 2 * s->ppu_2[jj * nh + ii] + c->gama_zmp / 2 * ppy;
                    G[jj * n + ii] = 2 * wx;
                    G[(nh + jj) * n + (nh + ii)] = 2 * wy;
                }
            /* cpp:517-526 gradient */
            double detpx[ORC_BODY_NH_MAX], detpy[ORC_BODY_NH_MAX];
            for (int k = 0; k < nh; k++) { detpx[k] = zx[k] - copx[k]; detpy[k] = zy[k] - copy_[k]; }
            double g0[2 * ORC_BODY_NH_MAX];
            {
                double S1[ORC_BODY_NH_MAX * ORC_BODY_NH_MAX], S2[ORC_BODY_NH_MAX * ORC_BODY_NH_MAX];
                double M1[ORC_BODY_NH_MAX * 2], M2[ORC_BODY_NH_MAX * 2];
                double t1[ORC_BODY_NH_MAX], t2[ORC_BODY_NH_MAX], t3[ORC_BODY_NH_MAX];
                /* alpha*pvu', beta*ppu' */
                for (int jj = 0; jj < nh; jj++) for (int ii = 0; ii < nh; ii++) {
                    S1[jj * nh + ii] = c->alphatheta * s->pvu[ii * nh + jj];
                    S2[jj * nh + ii] = c->beltatheta * s->ppu[ii * nh + jj];
                }
                mm(S1, s->pvs, M1, nh, nh, 2);
                mm(S2, s->pps, M2, nh, nh, 2);
                for (int half = 0; half < 2; half++) {
                    const double *th = half ? s->thetayk : s->thetaxk;
                    const double *bref = half ? by : bx;
                    mm(M1, th, t1, nh, 2, 1);
                    mm(M2, th, t2, nh, 2, 1);
                    mm(S2, bref, t3, nh, nh, 1);
                    for (int k = 0; k < nh; k++) {
                        /* gama * ptheta' * det: diagonal => exactly (gama*p_k) * det_k */
                        double w = half ? (c->gama_zmp * (-pth[k])) * detpx[k]
                                        : (c->gama_zmp * pth[k]) * detpy[k];
                        g0[half * nh + k] = ((t1[k] + t2[k]) - t3[k]) + w;
                    }
                }
            }
            /* cpp:542-561, 805-829: CI = -[q_upx q_lowx q_upy q_lowy t_upx t_lowx t_upy t_lowy]', rest zero */
            static __thread double CIbuf[2 * ORC_BODY_NH_MAX * 12 * ORC_BODY_NH_MAX];
            double ci0[12 * ORC_BODY_NH_MAX];
            double *CI = CIbuf;
            memset(CI, 0, sizeof(double) * n * m);
            memset(ci0, 0, sizeof(double) * m);
            double ppsx[ORC_BODY_NH_MAX], ppsy[ORC_BODY_NH_MAX];
            mm(s->pps, s->thetaxk, ppsx, nh, 2, 1);
            mm(s->pps, s->thetayk, ppsy, nh, 2, 1);
            for (int k = 0; k < nh; k++) {
                for (int jj = 0; jj < nh; jj++) {
                    double p = s->ppu[jj * nh + k];                     /* ppu(k,jj) */
                    CI[(0 * nh + k) * n + jj] = p * (-1);               /* -(q_upx)'  */
                    CI[(1 * nh + k) * n + jj] = (-p) * (-1);            /* -(q_lowx)' */
                    CI[(2 * nh + k) * n + nh + jj] = p * (-1);
                    CI[(3 * nh + k) * n + nh + jj] = (-p) * (-1);
                }
                CI[(4 * nh + k) * n + k] = c->j_ini * (-1);
                CI[(5 * nh + k) * n + k] = (-c->j_ini) * (-1);
                CI[(6 * nh + k) * n + nh + k] = c->j_ini * (-1);
                CI[(7 * nh + k) * n + nh + k] = (-c->j_ini) * (-1);
                ci0[0 * nh + k] = thmax - ppsx[k];
                ci0[1 * nh + k] = -thmin + ppsx[k];
                ci0[2 * nh + k] = thmax - ppsy[k];
                ci0[3 * nh + k] = -thmin + ppsy[k];
                ci0[4 * nh + k] = tqmax;
                ci0[5 * nh + k] = -tqmin;
                ci0[6 * nh + k] = tqmax;
                ci0[7 * nh + k] = -tqmin;
            }
            /* cpp:801-803, 837-849 */
            double X[2 * ORC_BODY_NH_MAX];
            memcpy(X, s->V_ini, sizeof(double) * n);
            s->status = orc_qp_solve(n, 0, m, G, g0, NULL, NULL, CI, ci0, X, &s->cost,
                                     s->active, &s->nactive, s->iters);
            s->qp_solution = 1;
            for (int k = 0; k < n; k++) if (isnan(X[k])) { s->qp_solution = 0; break; }
            memcpy(s->V_ini, X, sizeof(double) * n);

            /* cpp:567-625 first control, fallback / clamp */
            double ax0 = s->V_ini[0], ay0 = s->V_ini[nh];
            double arow_x = a[0] * s->thetaxk[0] + a[2] * s->thetaxk[1];   /* _a.row(0) * thetaxk */
            double arow_y = a[0] * s->thetayk[0] + a[2] * s->thetayk[1];
            if (!s->qp_solution) {
                ax0 = (s->thetaxk[0] - arow_x) / b[0];
                ay0 = (s->thetayk[0] - arow_y) / b[0];
            } else {
                double nx0 = arow_x + b[0] * ax0;
                if (nx0 > thmax) ax0 = (thmax - arow_x) / b[0];
                else if (nx0 < thmin) ax0 = (thmin - arow_x) / b[0];
                double ny0 = arow_y + b[0] * ay0;
                if (ny0 > thmax) ay0 = (thmax - arow_y) / b[0];
                else if (ny0 < thmin) ay0 = (thmin - arow_y) / b[0];
            }
            s->V_ini[0] = ax0;
            s->V_ini[nh] = ay0;
            double tmpx[2], tmpy[2];
            tmpx[0] = (a[0] * s->thetaxk[0] + a[2] * s->thetaxk[1]) + b[0] * ax0;
            tmpx[1] = (a[1] * s->thetaxk[0] + a[3] * s->thetaxk[1]) + b[1] * ax0;
            tmpy[0] = (a[0] * s->thetayk[0] + a[2] * s->thetayk[1]) + b[0] * ay0;
            tmpy[1] = (a[1] * s->thetayk[0] + a[3] * s->thetayk[1]) + b[1] * ay0;
            s->torquex_real0 = c->j_ini * ax0;
            s->torquey_real0 = c->j_ini * ay0;
            /* cpp:636-655 roll-out over the horizon */
            double xk[2] = { s->thetaxk[0], s->thetaxk[1] }, yk[2] = { s->thetayk[0], s->thetayk[1] };
            for (int jj = 0; jj < nh; jj++) {
                double ax = s->V_ini[jj], ay = s->V_ini[nh + jj], t0, t1_;
                t0 = (a[0] * xk[0] + a[2] * xk[1]) + b[0] * ax;
                t1_ = (a[1] * xk[0] + a[3] * xk[1]) + b[1] * ax;
                xk[0] = t0; xk[1] = t1_;
                s->thetax[jj] = xk[0];
                t0 = (a[0] * yk[0] + a[2] * yk[1]) + b[0] * ay;
                t1_ = (a[1] * yk[0] + a[3] * yk[1]) + b[1] * ay;
                yk[0] = t0; yk[1] = t1_;
                s->thetay[jj] = yk[0];
                s->zmpx_real[jj] = zx[jj] - c->j_ini * ay / (c->mass * (c->g + comacc_z_ref[jj]));
                s->zmpy_real[jj] = zy[jj] + c->j_ini * ax / (c->mass * (c->g + comacc_z_ref[jj]));
            }
            /* cpp:659-692 state advance + (normally zero-gain) feedback blend */
            s->thetaxk[0] = c->lamda[0] * bodyangle_state[0] + (1 - c->lamda[0]) * tmpx[0];
            s->thetaxk[1] = c->lamda[1] * bodyangle_state[1] + (1 - c->lamda[1]) * tmpx[1];
            s->thetayk[0] = c->lamda[2] * bodyangle_state[2] + (1 - c->lamda[2]) * tmpy[0];
            s->thetayk[1] = c->lamda[3] * bodyangle_state[3] + (1 - c->lamda[3]) * tmpy[1];
        }
    }
    /* cpp:696-709 (reads columns 0..2: needs nh >= 3) */
    out14[0] = s->thetax[0];     out14[1] = s->thetay[0];
    out14[2] = s->torquex_real0; out14[3] = s->torquey_real0;
    out14[4] = s->zmpx_real[0];  out14[5] = s->zmpy_real[0];
    out14[6] = s->thetax[1];     out14[7] = s->thetay[1];
    out14[8] = s->zmpx_real[1];  out14[9] = s->zmpy_real[1];
    out14[10] = s->thetax[2];    out14[11] = s->thetay[2];
    out14[12] = s->zmpx_real[2]; out14[13] = s->zmpy_real[2];
}

/*
 * Flat batch driver for the CPU baseline and the parity tests.  Layout = the
 * C-ABI's body-MPC record layout (include/go1mpc.h):
 *   tick[B]; tx[B][27]; theta_state[B][4] = (thetaxk0, thetaxk1, thetayk0, thetayk1) in/out;
 *   bodyangle_state[B][4]; refs[B][9*nh] = zmp x,y | bodyangle x,y | rfoot x,y | lfoot x,y | comacc_z;
 *   out14[B][14] in/out (a gated tick leaves it unchanged); x_out[B][2nh] in/out (V_ini).
 */
void orc_body_step_batch(const orc_body_cfg *c, int B, const int *tick,
                         const double *tx, double *theta_state,
                         const double *bodyangle_state,
                         const double *refs, double *out14,
                         double *x_out, int *active, int *nactive, int *iters, int *status)
{
    static __thread orc_body_mpc S;
    const int nh = c->nh;
    orc_body_init(&S, c);
    for (int bi = 0; bi < B; bi++) {
        const double *r = refs + (size_t)bi * 9 * nh;
        memcpy(S.tx, tx + (size_t)bi * ORC_FOOTSTEPS, sizeof S.tx);
        S.thetaxk[0] = theta_state[bi * 4 + 0]; S.thetaxk[1] = theta_state[bi * 4 + 1];
        S.thetayk[0] = theta_state[bi * 4 + 2]; S.thetayk[1] = theta_state[bi * 4 + 3];
        memcpy(S.V_ini, x_out + (size_t)bi * 2 * nh, sizeof(double) * 2 * nh);
        /* stale members = previous outputs */
        double *o = out14 + (size_t)bi * 14;
        S.thetax[0] = o[0]; S.thetay[0] = o[1]; S.torquex_real0 = o[2]; S.torquey_real0 = o[3];
        S.zmpx_real[0] = o[4]; S.zmpy_real[0] = o[5]; S.thetax[1] = o[6]; S.thetay[1] = o[7];
        S.zmpx_real[1] = o[8]; S.zmpy_real[1] = o[9]; S.thetax[2] = o[10]; S.thetay[2] = o[11];
        S.zmpx_real[2] = o[12]; S.zmpy_real[2] = o[13];
        S.status = -1; S.nactive = 0; memset(S.iters, 0, sizeof S.iters);
        orc_body_theta_mpc(&S, tick[bi], bodyangle_state + bi * 4, r, r + 2 * nh, r + 4 * nh,
                           r + 6 * nh, r + 8 * nh, o);
        theta_state[bi * 4 + 0] = S.thetaxk[0]; theta_state[bi * 4 + 1] = S.thetaxk[1];
        theta_state[bi * 4 + 2] = S.thetayk[0]; theta_state[bi * 4 + 3] = S.thetayk[1];
        memcpy(x_out + (size_t)bi * 2 * nh, S.V_ini, sizeof(double) * 2 * nh);
        if (active) memcpy(active + (size_t)bi * 12 * nh, S.active, sizeof(int) * (S.nactive > 0 ? S.nactive : 0));
        if (nactive) nactive[bi] = S.nactive;
        if (iters) memcpy(iters + bi * 4, S.iters, sizeof S.iters);
        if (status) status[bi] = S.status;
    }
}
