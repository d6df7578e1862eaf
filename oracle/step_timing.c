/*
 * step_timing.c -- oracle restatement of the step-location / step-timing SQP tick.
 *
 * TEST INFRASTRUCTURE (see go1_oracle.h).  Restates NLPClass (NLP = unitree_ros/mosek_nlp_kmp):
 *   step_timing_opti_loop        NLP/src/NLP/NLPClass_sqp.cpp:693-1102
 *   Indexfind                    :1105-1141
 *   step_timing_object_function  :1144-1173
 *   step_timing_constraints      :1175-1458
 *   solve_stepping_timing/Solve  :1613-1653
 * Every expression keeps the association the reference's Eigen expression has when its
 * products are evaluated in ascending inner index (what oracle/eigen_shim does), so this file
 * and the unmodified NLPClass compiled against the shim agree bit for bit
 * (tests/test_oracle_vs_ref.py).  Products with the selection rows _SS1.._SS4 (unit rows)
 * reduce exactly to picking an entry: the other terms are exact zeros.
 *
 *   CoM_height_solve             :2361-2473 (inverse of the 7x7 time-polynomial matrix by
 *                                row-pivoted Gauss-Jordan, first maximal pivot wins -- what
 *                                oracle/eigen_shim's inverse() does for n > 3)
 * Frozen quirks: `_vari_ini += _X` runs whatever the solver's status was (a not-PD solve
 * leaves _X = _vari_ini, doubling it; an infeasible solve adds a partial step); the
 * initial-velocity rows 20-23 use a.dt in the matrix and a.dt/2 in the right-hand side; the
 * swing-velocity rows 8-11 are all-zero rows with b = 0 while k_yu == 0; the feedback blend
 * computes ((1-l)(com - p) + l est) + p even with l = 0 (one rounding).
 * Differences: Indexfind's unbounded while loops are clamped to the 27-entry table.
 */
#include <math.h>
#include <string.h>
#include "go1_oracle.h"

#define NS ORC_FOOTSTEPS

void orc_step_cfg_default(orc_step_cfg *c)
{
    /* NLPClass_sqp.cpp:212-214,244-258,273-286 ; NLPRTControlClass.cpp:35-42 ; NLPClass.h:30-38 */
    memset(c, 0, sizeof *c);
    c->dt = 0.025;
    c->ggg = 9.8;
    c->Wn = sqrt(9.8 / (0.309458 - 0.000));
    c->t_min = 0.5; c->t_max = 1;
    c->footx_max = 0.15; c->footx_min = -0.05;
    c->footx_vmax = 3; c->footx_vmin = -2.875; c->footy_vmax = 2; c->footy_vmin = -1;
    c->comax_max = 5; c->comax_min = -5; c->comay_max = 6; c->comay_min = -6;
    c->aax = 50000; c->aay = 50000; c->aaxv = 1000; c->aayv = 500;
    c->bbx = 2000000; c->bby = 10000000; c->rr1 = 1000000; c->rr2 = 1000000;
    c->half_hip_width = 0.12675; c->foot_width = 0.03;
    c->hcom = 0.309458 - 0.000;
    c->n_sqp = 3;
}

void orc_step_state_default(orc_step_state *s, const orc_step_cfg *c, double steplength, double stepwidth,
                            double stepheight, double tstep)
{
    /* FootStepInputs :51-75, Initialize :131-206 */
    double sl[NS], sw[NS], sh[NS];
    memset(s, 0, sizeof *s);
    for (int j = 0; j < NS; j++) { sl[j] = steplength; sw[j] = stepwidth; sh[j] = stepheight; }
    sl[NS - 1] = sl[NS - 2] = sl[NS - 3] = sl[NS - 4] = sl[NS - 5] = 0;
    sl[0] = sl[1] = sl[2] = 0; sl[3] = steplength / 2;
    sw[0] = sw[0] / 2;
    sl[14] = 0;
    for (int j = 15; j <= 21; j++) sl[j] *= -1;
    for (int j = 0; j < NS; j++) { s->Lxx[j] = sl[j]; s->Lyy[j] = (int)pow(-1, j) * sw[j]; }
    for (int j = 1; j < NS; j++) {
        s->footx[j] = s->footx[j - 1] + sl[j - 1];
        s->footy[j] = s->footy[j - 1] + (int)pow(-1, j - 1) * sw[j - 1];
        s->footz[j] = s->footz[j - 1] + sh[j - 1];
    }
    for (int j = 0; j < NS; j++) s->ts[j] = tstep;
    for (int j = 1; j < NS; j++) {
        s->tx[j] = s->tx[j - 1] + s->ts[j - 1];
        s->tx[j] = round(s->tx[j] / c->dt) * c->dt - 0.000001;
    }
}

/* inverse by Gauss-Jordan with partial (row) pivoting, first maximal |pivot| wins; row-major n x n */
static void gj_inverse(int n, double *a, double *r)
{
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
        double d = a[k * n + k];
        for (int j = 0; j < n; j++) { a[k * n + j] = a[k * n + j] / d; r[k * n + j] = r[k * n + j] / d; }
        for (int i = 0; i < n; i++) {
            if (i == k) continue;
            double f = a[i * n + k];
            for (int j = 0; j < n; j++) { a[i * n + j] -= f * a[k * n + j]; r[i * n + j] -= f * r[k * n + j]; }
        }
    }
}

/* NLPClass::CoM_height_solve :2361-2473 for the three samples the tick reads (jxx = 1..3) */
static void com_height_solve(const orc_step_cfg *c, int i, const orc_step_state *s, int bjx1, int nsamp,
                             double *comz, double *comvz, double *comaz)
{
    if (bjx1 >= 2) {
        double tp[3] = { 0.0001, s->ts[bjx1 - 1] / 2 + 0.0001, s->ts[bjx1 - 1] + 0.0001 };
        double A[49], Ainv[49];
        const int rowt[7] = { 0, 0, 0, 1, 2, 2, 2 }, kind[7] = { 1, 2, 0, 0, 0, 1, 2 };   /* 0 value, 1 velocity, 2 acceleration */
        for (int r = 0; r < 7; r++) {
            const double t = tp[rowt[r]];
            double *a = A + 7 * r;
            if (kind[r] == 0) { a[0] = pow(t, 6); a[1] = pow(t, 5); a[2] = pow(t, 4); a[3] = pow(t, 3); a[4] = pow(t, 2); a[5] = pow(t, 1); a[6] = 1; }
            else if (kind[r] == 1) { a[0] = 6 * pow(t, 5); a[1] = 5 * pow(t, 4); a[2] = 4 * pow(t, 3); a[3] = 3 * pow(t, 2); a[4] = 2 * pow(t, 1); a[5] = 1; a[6] = 0; }
            else { a[0] = 30 * pow(t, 4); a[1] = 20 * pow(t, 3); a[2] = 12 * pow(t, 2); a[3] = 6 * pow(t, 1); a[4] = 2; a[5] = 0; a[6] = 0; }
        }
        gj_inverse(7, A, Ainv);
        const double f0 = s->footz[bjx1 - 2], f1 = s->footz[bjx1 - 1];
        const double plan[7] = { 0, 0, f0 + c->hcom, (f0 + f1) / 2 + c->hcom, f1 + c->hcom, 0, 0 };
        double co[7];
        for (int r = 0; r < 7; r++) { double acc = 0.0; for (int k = 0; k < 7; k++) acc += Ainv[7 * r + k] * plan[k]; co[r] = acc; }
        for (int jxx = 1; jxx <= nsamp; jxx++) {
            const double t = (i + jxx - round(s->tx[bjx1 - 1] / c->dt)) * c->dt;
            const double p[7] = { pow(t, 6), pow(t, 5), pow(t, 4), pow(t, 3), pow(t, 2), pow(t, 1), 1 };
            const double v[7] = { 6 * pow(t, 5), 5 * pow(t, 4), 4 * pow(t, 3), 3 * pow(t, 2), 2 * pow(t, 1), 1, 0 };
            const double a[7] = { 30 * pow(t, 4), 20 * pow(t, 3), 12 * pow(t, 2), 6 * pow(t, 1), 2, 0, 0 };
            double z = 0.0, vz = 0.0, az = 0.0;
            for (int k = 0; k < 7; k++) { z += p[k] * co[k]; vz += v[k] * co[k]; az += a[k] * co[k]; }
            comz[jxx - 1] = z; comvz[jxx - 1] = vz; comaz[jxx - 1] = az;
        }
    } else {
        /* the reference writes three samples; beyond them the arrays keep their initial values (_hcom, 0, 0) */
        for (int q = 0; q < nsamp; q++) { comz[q] = c->hcom; comvz[q] = 0; comaz[q] = 0; }
    }
}

void orc_step_timing_tick(const orc_step_cfg *c, int i, orc_step_state *s, const orc_step_in *in,
                          double out38[38], orc_step_diag *dg)
{
    orc_step_timing_tick_ext(c, i, s, in, out38, dg, NULL);
}

void orc_step_timing_tick_ext(const orc_step_cfg *c, int i, orc_step_state *s, const orc_step_in *in,
                              double out38[38], orc_step_diag *dg, orc_step_ext *ext)
{
    const double dt = c->dt, Wn = c->Wn;
    /* :933 _nTdx = round(_td(1) / _dt) + 1 with _td = 0.2 _ts as the END of the previous tick left it (:1041, :201) */
    int ntdx = (int)round(0.2 * s->ts[1] / dt) + 1;
    if (ntdx > ORC_NTD_MAX) ntdx = ORC_NTD_MAX;
    if (ntdx < 3 || !ext) ntdx = 3;
    double v[4];
    if (dg) memset(dg, 0, sizeof *dg);

    /* :702-704  Indexfind((i+1)*dt, xyz0 = -1) */
    int j = 0;
    while (j < NS && (i + 1) * dt > s->tx[j] + 0.0001) j++;
    const int p = (j - 1) + 1;                      /* _periond_i */
    const double px = s->footx[p - 1], py = s->footy[p - 1];
    /* :714-727 */
    const int ki = (int)round(s->tx[p - 1] / dt);
    const int k_yu = i - ki;
    const double Tk = s->ts[p - 1] - k_yu * dt;
    const double Lxx_refx = s->Lxx[p - 1], Lyy_refy = s->Lyy[p - 1];
    const double tr1_ref = cosh(Wn * Tk), tr2_ref = sinh(Wn * Tk);
    /* :730-740 warm start */
    if (i == 1) { v[0] = Lxx_refx; v[1] = Lyy_refy; v[2] = tr1_ref; v[3] = tr2_ref; }
    else memcpy(v, s->vari, sizeof v);
    /* :745-757 remaining-time bounds */
    double tr1_min, tr2_min;
    if ((c->t_min - k_yu * dt) >= 0.001) { tr1_min = cosh(Wn * (c->t_min - k_yu * dt)); tr2_min = sinh(Wn * (c->t_min - k_yu * dt)); }
    else { tr1_min = cosh(Wn * (0.001)); tr2_min = sinh(Wn * (0.001)); }
    const double tr1_max = cosh(Wn * (c->t_max - k_yu * dt)), tr2_max = sinh(Wn * (c->t_max - k_yu * dt));

    const double comx_f = s->feed[0], comvx_f = s->feed[1], comy_f = s->feed[3], comvy_f = s->feed[4];
    double endx = s->endref[0], endy = s->endref[1];
    if (i == 1) {
        /* :761-771 */
        double isx = comx_f - px, esx = v[0] * 0.5, visx = (esx - isx * v[2]) / (1 / Wn * v[3]);
        double isy = comy_f - py, esy = v[1] * 0.5, visy = (esy - isy * v[2]) / (1 / Wn * v[3]);
        endx = Wn * isx * v[3] + visx * v[2];
        endy = Wn * isy * v[3] + visy * v[2];
    }

    /* quantities of the objective / constraints that do not change inside the SQP loop */
    const double AxO = comx_f - px, BxO = comvx_f / Wn, Cx = -0.5 * Lxx_refx;
    const double Axv = Wn * BxO, Bxv = Wn * AxO, Cxv = -endx;
    const double AyO = comy_f - py, ByO = comvy_f / Wn, Cy = -0.5 * Lyy_refy;
    const double Ayv = Wn * ByO, Byv = Wn * AyO, Cyv = -endy;
    const double aax = c->aax, aay = c->aay, aaxv = c->aaxv, aayv = c->aayv;
    double SQ0[4][4];
    memset(SQ0, 0, sizeof SQ0);
    SQ0[0][0] = 0.5 * c->bbx;
    SQ0[1][1] = 0.5 * c->bby;
    SQ0[2][2] = 0.5 * (c->rr1 + aax * AxO * AxO + aay * AyO * AyO + aaxv * Axv * Axv + aayv * Ayv * Ayv);
    SQ0[2][3] = 0.5 * (aax * AxO * BxO + aay * AyO * ByO + aaxv * Axv * Bxv + aayv * Ayv * Byv);
    SQ0[3][2] = 0.5 * (aax * BxO * AxO + aay * ByO * AyO + aaxv * Bxv * Axv + aayv * Byv * Ayv);
    SQ0[3][3] = 0.5 * (c->rr2 + aax * BxO * BxO + aay * ByO * ByO + aaxv * Bxv * Bxv + aayv * Byv * Byv);
    double SQ[4][4], Sq[4];
    for (int r = 0; r < 4; r++) for (int k = 0; k < 4; k++) SQ[r][k] = (SQ0[r][k] + SQ0[k][r]) / 2.0;
    Sq[0] = -c->bbx * Lxx_refx;
    Sq[1] = -c->bby * Lyy_refy;
    Sq[2] = -c->rr1 * tr1_ref + aax * AxO * Cx + aay * AyO * Cy + aaxv * Axv * Cxv + aayv * Ayv * Cyv;
    Sq[3] = -c->rr2 * tr2_ref + aax * BxO * Cx + aay * ByO * Cy + aaxv * Bxv * Cxv + aayv * Byv * Cyv;

    /* :1216-1246 lateral reachability, widened after two steps */
    double footy_max, footy_min;
    const int wide = (i >= (round(2 * s->ts[1] / dt)) + 1);
    const double HW = c->half_hip_width, FW = c->foot_width;
    if (p % 2 == 0) { footy_min = -(2 * HW + 0.03); footy_max = wide ? -(FW + 0.01) : -(HW - 0.03); }
    else { footy_max = 2 * HW + 0.03; footy_min = wide ? FW + 0.01 : HW - 0.03; }

    /* :1341-1411 coefficient rows that depend on the state only */
    const double CCx = comx_f - px, CCy = comy_f - py;
    const double AA = Wn * sinh(Wn * dt);
    const double BBx = pow(Wn, 2) * CCx * cosh(Wn * dt), BBy = pow(Wn, 2) * CCy * cosh(Wn * dt);
    const double AA1x = AA * Wn, AA2x = -2 * AA * CCx * Wn, AA3x = 2 * BBx;
    const double AA1y = AA * Wn, AA2y = -2 * AA * CCy * Wn, AA3y = 2 * BBy;
    const double VAA = cosh(Wn * dt);
    const double VBBx = Wn * CCx * sinh(Wn * dt), VBBy = Wn * CCy * sinh(Wn * dt);
    const double VAA1x = VAA * Wn, VAA2x = -2 * VAA * CCx * Wn, VAA3x = 2 * VBBx - 2 * comvx_f;
    const double VAA1y = VAA * Wn, VAA2y = -2 * VAA * CCy * Wn, VAA3y = 2 * VBBy - 2 * comvy_f;
    const double VAA1x1 = Wn, VAA2x1 = -2 * CCx * Wn, VAA3x1 = -2 * comvx_f;
    const double VAA1y1 = Wn, VAA2y1 = -2 * CCy * Wn, VAA3y1 = -2 * comvy_f;

    int n_solved = 0;
    for (int it = 1; it <= c->n_sqp; it++) {
        /* :1164-1165 */
        double G[16], g0[4];
        for (int r = 0; r < 4; r++) for (int k = 0; k < 4; k++) G[k * 4 + r] = 2 * SQ[r][k];
        for (int r = 0; r < 4; r++) {
            double acc = 0.0;
            for (int k = 0; k < 4; k++) acc += (2 * SQ[r][k]) * v[k];
            g0[r] = acc + Sq[r];
        }
        /* :1188-1190 linearised tr1^2 - tr2^2 = 1 */
        double CE[4], ce0[1];
        {
            double trx12[4] = { 0.0, 0.0, 2 * v[2], (-2) * v[3] };
            double q = 0.0;
            q += v[2] * v[2];
            q += (v[3] * (-1)) * v[3];
            ce0[0] = -q + 1;
            for (int k = 0; k < 4; k++) CE[k] = trx12[k] * (-1);
        }
        /* rows A x <= b */
        double A[24][4], b[24];
        memset(A, 0, sizeof A);
        memset(b, 0, sizeof b);
        A[0][2] = 1;  b[0] = -(v[2]) + tr1_max;
        A[1][2] = -1; b[1] = -((-1.0) * v[2]) - tr1_min;
        A[2][3] = 1;  b[2] = -(v[3]) + tr2_max;
        A[3][3] = -1; b[3] = -((-1.0) * v[3]) - tr2_min;
        A[4][0] = 1;  b[4] = -(v[0]) + c->footx_max;
        A[5][0] = -1; b[5] = -((-1.0) * v[0]) - c->footx_min;
        A[6][1] = 1;  b[6] = -(v[1]) + footy_max;
        A[7][1] = -1; b[7] = -((-1.0) * v[1]) - footy_min;
        if (k_yu != 0) {
            A[8][0] = 1;   b[8] = -(v[0] - s->Lxx[p - 1] - c->footx_vmax * dt);
            A[9][0] = -1;  b[9] = v[0] - s->Lxx[p - 1] - c->footx_vmin * dt;
            A[10][1] = 1;  b[10] = -(v[1] - s->Lyy[p - 1] - c->footy_vmax * dt);
            A[11][1] = -1; b[11] = v[1] - s->Lyy[p - 1] - c->footy_vmin * dt;
        }
#define ROW3(r, i0, c0, c2, c3, d3) do { \
            A[r][i0] = (c0); A[r][2] = (c2); A[r][3] = (c3); \
            double acc_ = 0.0; \
            if ((i0) == 0) { acc_ += (-(c0)) * v[0]; } else { acc_ += (-(c0)) * v[1]; } \
            acc_ += (-(c2)) * v[2]; acc_ += (-(d3)) * v[3]; b[r] = acc_; } while (0)
        {
            double c3;
            /* CoM acceleration at the next sample :1349-1363 */
            c3 = AA3x - 2 * c->comax_max; ROW3(12, 0, AA1x, AA2x, c3, c3);
            c3 = -(AA3x - 2 * c->comax_min); ROW3(13, 0, -AA1x, -AA2x, c3, c3);
            c3 = AA3y - 2 * c->comay_max; ROW3(14, 1, AA1y, AA2y, c3, c3);
            c3 = -(AA3y - 2 * c->comay_min); ROW3(15, 1, -AA1y, -AA2y, c3, c3);
            /* CoM velocity increment over one dt :1374-1387 */
            c3 = VAA3x - 2 * c->comax_max * dt; ROW3(16, 0, VAA1x, VAA2x, c3, c3);
            c3 = -(VAA3x - 2 * c->comax_min * dt); ROW3(17, 0, -VAA1x, -VAA2x, c3, c3);
            c3 = VAA3y - 2 * c->comay_max * dt; ROW3(18, 1, VAA1y, VAA2y, c3, c3);
            c3 = -(VAA3y - 2 * c->comay_min * dt); ROW3(19, 1, -VAA1y, -VAA2y, c3, c3);
            /* CoM initial velocity :1397-1411 -- matrix uses a dt, right-hand side a dt / 2 */
            double d3;
            c3 = VAA3x1 - 2 * c->comax_max * dt; d3 = VAA3x1 - 2 * c->comax_max * dt / 2.0; ROW3(20, 0, VAA1x1, VAA2x1, c3, d3);
            c3 = -(VAA3x1 - 2 * c->comax_min * dt); d3 = -(VAA3x1 - 2 * c->comax_min * dt / 2.0); ROW3(21, 0, -VAA1x1, -VAA2x1, c3, d3);
            c3 = VAA3y1 - 2 * c->comay_max * dt; d3 = VAA3y1 - 2 * c->comay_max * dt / 2.0; ROW3(22, 1, VAA1y1, VAA2y1, c3, d3);
            c3 = -(VAA3y1 - 2 * c->comay_min * dt); d3 = -(VAA3y1 - 2 * c->comay_min * dt / 2.0); ROW3(23, 1, -VAA1y1, -VAA2y1, c3, d3);
        }
#undef ROW3
        if (Tk >= 0.1 * s->ts[p - 1]) {
            /* :1618-1638 */
            double CI[96], X[4], cost;
            for (int r = 0; r < 24; r++) for (int k = 0; k < 4; k++) CI[r * 4 + k] = A[r][k] * (-1);
            memcpy(X, v, sizeof X);
            int act[26], na = 0, iters[4] = { 0, 0, 0, 0 };
            int st = orc_qp_solve(4, 1, 24, G, g0, CE, ce0, CI, b, X, &cost, act, &na, iters);
            if (dg && n_solved < ORC_STEP_NQP_MAX) {
                dg->status[n_solved] = st; dg->nactive[n_solved] = na;
                memcpy(dg->iters[n_solved], iters, sizeof iters);
                memcpy(dg->active[n_solved], act, sizeof(int) * (na < 25 ? na : 25));
                memcpy(dg->x[n_solved], X, sizeof X);
            }
            n_solved++;
            for (int k = 0; k < 4; k++) v[k] += X[k];    /* :795-798, whatever the status */
        } else {
            v[0] = Lxx_refx; v[1] = Lyy_refy; v[2] = tr1_ref; v[3] = tr2_ref;
        }
    }

    /* :817, :886-888 write-back */
    memcpy(s->vari, v, sizeof v);
    s->Lxx[p - 1] = v[0];
    s->Lyy[p - 1] = v[1];
    s->ts[p - 1] = k_yu * dt + log(v[2] + v[3]) / Wn;
    /* :896-901 */
    const double isx = comx_f - px, esx = v[0] * 0.5, visx = (esx - isx * v[2]) / (1 / Wn * v[3]);
    const double isy = comy_f - py, esy = v[1] * 0.5, visy = (esy - isy * v[2]) / (1 / Wn * v[3]);
    /* :906-913 */
    const double nTd_ts1 = s->ts[1];    /* NB: _td = 0.2*_ts was taken at the END of the previous tick; ts[1] only changes while p-1 == 1 */
    for (int jxx = p + 1; jxx <= NS; jxx++) s->tx[jxx - 1] = s->tx[jxx - 2] + s->ts[jxx - 2];
    if (p < NS) { s->footx[p] = s->footx[p - 1] + v[0]; s->footy[p] = s->footy[p - 1] + v[1]; }
    s->endref[0] = Wn * isx * v[3] + visx * v[2];
    s->endref[1] = Wn * isy * v[3] + visy * v[2];
    (void)nTd_ts1;

    /* :936 vertical CoM samples (after the write-back of ts / tx, with the _bjx1 of the previous tick) */
    double hz_z[ORC_NTD_MAX], hz_vz[ORC_NTD_MAX], hz_az[ORC_NTD_MAX];
    if (c->ext_height) {
        for (int q = 0; q < 3; q++) { hz_z[q] = in->comz[q]; hz_az[q] = in->comaz[q]; hz_vz[q] = 0.0; }
        hz_vz[0] = in->comvz0;
        for (int q = 3; q < ntdx; q++) { hz_z[q] = in->comz[2]; hz_az[q] = in->comaz[2]; hz_vz[q] = 0.0; }   /* external heights: held */
    } else {
        com_height_solve(c, i, s, (int)s->bjx1_prev, ntdx, hz_z, hz_vz, hz_az);
    }
    /* :938-955 LIPM roll-out (samples i, i+1, i+2 are what the outputs read) */
    double comx[3], comy[3], comvx[3], comvy[3], comax[3], comay[3], zmpx[3], zmpy[3], dcmx[3], dcmy[3];
    for (int jxx = 1; jxx <= 3; jxx++) {
        const int q = jxx - 1;
        const double w = Wn * dt * jxx;
        comx[q] = isx * cosh(w) + visx * 1 / Wn * sinh(w) + px;
        comy[q] = isy * cosh(w) + visy * 1 / Wn * sinh(w) + py;
        comvx[q] = Wn * isx * sinh(w) + visx * cosh(w);
        comvy[q] = Wn * isy * sinh(w) + visy * cosh(w);
        comax[q] = pow(Wn, 2) * isx * cosh(w) + visx * Wn * sinh(w);
        comay[q] = pow(Wn, 2) * isy * cosh(w) + visy * Wn * sinh(w);
        const double hz = (hz_z[q] - in->zsc[q]) / (hz_az[q] + c->ggg);
        zmpx[q] = comx[q] - hz * comax[q];
        zmpy[q] = comy[q] - hz * comay[q];
        dcmx[q] = comx[q] + comvx[q] * sqrt(hz);
        dcmy[q] = comy[q] + comvy[q] * sqrt(hz);
    }
    if (ext) {
        /* the roll-out runs to jxx = _nTdx (:938); only NLPClass::Zmp_distributor ever reads samples beyond i + 2 */
        ext->ntdx = ntdx;
        for (int jxx = 1; jxx <= ntdx; jxx++) {
            const int q = jxx - 1;
            if (q < 3) { ext->zmpx[q] = zmpx[q]; ext->zmpy[q] = zmpy[q]; continue; }
            const double w = Wn * dt * jxx;
            const double cx = isx * cosh(w) + visx * 1 / Wn * sinh(w) + px;
            const double cy = isy * cosh(w) + visy * 1 / Wn * sinh(w) + py;
            const double ax = pow(Wn, 2) * isx * cosh(w) + visx * Wn * sinh(w);
            const double ay = pow(Wn, 2) * isy * cosh(w) + visy * Wn * sinh(w);
            const double hz = (hz_z[q] - ext->zsc[q]) / (hz_az[q] + c->ggg);
            ext->zmpx[q] = cx - hz * ax;
            ext->zmpy[q] = cy - hz * ay;
        }
    }
    /* :963-972, :1017-1022 feedback blend (gains 0 as shipped) */
    double e0 = in->est[0], e3 = in->est[3];
    if (p % 2 == 0) { e0 = e0 - in->lfoot_fb[0]; e3 = e3 - in->lfoot_fb[1]; }
    else { e0 = e0 - in->rfoot_fb[0]; e3 = e3 - in->rfoot_fb[1]; }
    const double lx = c->lamda[0], lvx = c->lamda[1], ly = c->lamda[2], lvy = c->lamda[3];
    s->feed[0] = ((1 - lx) * (comx[0] - px) + (lx) * e0) + px;
    s->feed[1] = (1 - lvx) * comvx[0] + (lvx) * in->est[1];
    s->feed[2] = (1 - lx) * comax[0] + lx * in->est[2];
    s->feed[3] = ((1 - ly) * (comy[0] - py) + (ly) * e3) + py;
    s->feed[4] = (1 - lvy) * comvy[0] + (lvy) * in->est[4];
    s->feed[5] = (1 - ly) * comay[0] + ly * in->est[5];

    /* :1031-1041 integer step indices against the UPDATED table (xyz1 = 0 branch) */
    j = 0; while (j < NS && i * dt >= s->tx[j]) j++;
    const int bjxx = (j - 1) + 1;
    j = 0; while (j < NS && (i + 1) * dt >= s->tx[j]) j++;
    const int bjx1 = (j - 1) + 1;

    /* :1048-1090 */
    out38[0] = comx[0]; out38[1] = comy[0]; out38[2] = hz_z[0];
    out38[3] = comvx[0]; out38[4] = comvy[0]; out38[5] = hz_vz[0];
    out38[6] = comax[0]; out38[7] = comay[0]; out38[8] = hz_az[0];
    out38[9] = zmpx[0]; out38[10] = zmpy[0]; out38[11] = dcmx[0]; out38[12] = dcmy[0];
    out38[13] = zmpx[1]; out38[14] = zmpy[1]; out38[15] = dcmx[1]; out38[16] = dcmy[1];
    out38[17] = zmpx[2]; out38[18] = zmpy[2]; out38[19] = dcmx[2]; out38[20] = dcmy[2];
    out38[21] = comax[1]; out38[22] = comay[1]; out38[23] = hz_az[1];
    out38[24] = comax[2]; out38[25] = comay[2]; out38[26] = hz_az[2];
    out38[27] = bjxx;
    const int b0 = bjxx < NS ? bjxx : NS - 1, b1 = bjxx + 1 < NS ? bjxx + 1 : NS - 1;   /* reference reads past the table at the very end */
    out38[28] = s->footx[b0]; out38[29] = s->footx[b1];
    out38[30] = s->footy[b0]; out38[31] = s->footy[b1];
    out38[32] = s->footz[b0]; out38[33] = s->footz[b1];
    out38[34] = p - 1;
    out38[35] = s->ts[p - 1];
    out38[36] = v[0];
    out38[37] = v[1];
    s->bjx1_prev = bjx1;
    if (dg) { dg->periond_i = p; dg->k_yu = k_yu; dg->bjxx = bjxx; dg->bjx1 = bjx1; dg->n_solved = n_solved; }
}

void orc_step_timing_batch(const orc_step_cfg *c, int B, const int *tick, double *states, const double *ins,
                           double *out38, orc_step_diag *diag)
{
    for (int bi = 0; bi < B; bi++)
        orc_step_timing_tick(c, tick[bi], (orc_step_state *)(states + (size_t)bi * (sizeof(orc_step_state) / sizeof(double))),
                             (const orc_step_in *)(ins + (size_t)bi * (sizeof(orc_step_in) / sizeof(double))),
                             out38 + (size_t)bi * 38, diag ? diag + bi : NULL);
}

/* ===================== swing-foot trajectory (NLPClass::Foot_trajectory_solve_mod2) ===================== */
void orc_foot_state_default(double fs[ORC_FOOT_STATE], double sw0)
{
    /* NLPClass_sqp.cpp:496-500: _Rfooty = -stepwidth(0), _Lfooty = +stepwidth(0), the rest 0 */
    for (int k = 0; k < 4; k++) {
        double *p = fs + 6 * k;
        p[0] = 0; p[1] = -sw0; p[2] = 0; p[3] = 0; p[4] = sw0; p[5] = 0;
    }
    for (int k = 24; k < 30; k++) fs[k] = 0.0;
    fs[30] = -1.0;
    fs[31] = 0.0;
}

int orc_foot_traj_tick(const orc_step_cfg *c, int j, const orc_step_state *st, int bjxx,
                       double fs[ORC_FOOT_STATE], double sw0, double lift_height, double out18[18])
{
    const double dt = c->dt;
    const int bjx1 = (int)st->bjx1_prev;             /* _bjx1 as the tick just run left it */
    const int t_end = (int)round((st->tx[NS - 1] - 2 * 0.7) / dt);   /* :593 is evaluated at Initialize; see note */
    double *pm1 = fs, *pj = fs + 6, *pm2 = fs + 12, *pm3 = fs + 18, *frz = fs + 24;
    double cur[6], nxt[6], vel[6] = { 0, 0, 0, 0, 0, 0 }, acc[6] = { 0, 0, 0, 0, 0, 0 };
    int wrote_next[6] = { 0, 0, 0, 0, 0, 0 };
    int right_support;
    memcpy(cur, pj, sizeof cur);
    /* _footxyz_real: the step tables with (1,0) overwritten by -stepwidth(0) (:2050) */
#define FXR(r, k) ((r) == 0 ? st->footx[k] : ((r) == 1 ? ((k) == 0 ? -sw0 : st->footy[k]) : st->footz[k]))
    (void)t_end;
    if (bjx1 >= 2) {
        const int s = (int)round(st->tx[bjx1 - 1] / dt);
        if ((double)s != fs[30]) {
            /* first tick of this step: freeze the feet where they stood at index s-2 */
            const int back = j - (s - 2);            /* 1, 2 or 3 ticks ago */
            const double *src = (back <= 1) ? pm1 : (back == 2 ? pm2 : pm3);
            memcpy(frz, src, sizeof(double) * 6);
            fs[30] = (double)s;
        }
        const int left_support = (bjx1 % 2 == 0);
        const int so = left_support ? 3 : 0;         /* support foot offset in the 6-vectors (L : R) */
        const int wo = left_support ? 0 : 3;         /* swing foot offset                              */
        right_support = left_support ? 0 : 1;
        for (int k = 0; k < 3; k++) { cur[so + k] = frz[so + k]; nxt[so + k] = frz[so + k]; wrote_next[so + k] = 1; }
        if ((j + 1 - s) * dt < 0.2 * st->ts[bjx1 - 1]) {
            right_support = 2;
            for (int k = 0; k < 3; k++) { cur[wo + k] = frz[wo + k]; nxt[wo + k] = frz[wo + k]; wrote_next[wo + k] = 1; }
        } else {
            const double t_des = (j + 1 - s + 1) * dt;
            const double td1 = 0.2 * st->ts[bjx1 - 1], ts1 = st->ts[bjx1 - 1];
            const double tp[3] = { t_des - dt, (td1 + ts1) / 2 + 0.0001, ts1 };
            if (fabs(t_des - ts1) <= (+0.0005)) {
                for (int k = 0; k < 3; k++) { cur[wo + k] = FXR(k, bjxx); nxt[wo + k] = FXR(k, bjxx); wrote_next[wo + k] = 1; }
            } else {
                double A[16], Ai[16];
                for (int r = 0; r < 3; r++) { A[4 * r] = pow(tp[r], 3); A[4 * r + 1] = pow(tp[r], 2); A[4 * r + 2] = pow(tp[r], 1); A[4 * r + 3] = 1; }
                A[12] = 3 * pow(tp[2], 2); A[13] = 2 * pow(tp[2], 1); A[14] = pow(tp[2], 0); A[15] = 0;
                gj_inverse(4, A, Ai);
                const double tap[4] = { pow(t_des, 3), pow(t_des, 2), pow(t_des, 1), 1 };
                const double tav[4] = { 3 * pow(t_des, 2), 2 * pow(t_des, 1), 1, 0 };
                const double taa[4] = { 6 * pow(t_des, 1), 2, 0, 0 };
                if ((j + 1 - s) * dt < td1 + dt)
                    fs[31] = (FXR(1, bjxx) + FXR(1, bjxx - 2)) / 2;
                for (int k = 0; k < 3; k++) {
                    double plan[4];
                    plan[0] = pm1[wo + k];
                    if (k == 0) plan[1] = (FXR(0, bjxx - 2) + FXR(0, bjxx)) / 2;
                    else if (k == 1) plan[1] = fs[31];
                    else plan[1] = fmax(FXR(2, bjxx - 2), FXR(2, bjxx)) + ((bjx1 - 1 >= NS - 2) ? 0.0 : lift_height);
                    plan[2] = FXR(k, bjxx);
                    plan[3] = 0;
                    double co[4];
                    for (int r = 0; r < 4; r++) { double a_ = 0.0; for (int q = 0; q < 4; q++) a_ += Ai[4 * r + q] * plan[q]; co[r] = a_; }
                    double p_ = 0.0, v_ = 0.0, a2 = 0.0;
                    for (int q = 0; q < 4; q++) { p_ += tap[q] * co[q]; v_ += tav[q] * co[q]; a2 += taa[q] * co[q]; }
                    cur[wo + k] = p_; vel[wo + k] = v_; acc[wo + k] = a2;
                    nxt[wo + k] = cur[wo + k] + dt * vel[wo + k];
                    wrote_next[wo + k] = 1;
                }
            }
        }
    } else {
        right_support = 2;
        cur[1] = -sw0;      /* :2316-2317 only the lateral coordinates are set */
        cur[4] = sw0;
    }
#undef FXR
    /* :2320-2338 outputs: R xyz, L xyz | velocities | accelerations at tick j */
    for (int k = 0; k < 6; k++) { out18[k] = cur[k]; out18[6 + k] = vel[k]; out18[12 + k] = acc[k]; }
    /* slide the window: what the arrays hold at j+1 is this tick's extrapolation, or their initial value */
    memcpy(pm3, pm2, sizeof(double) * 6);
    memcpy(pm2, pm1, sizeof(double) * 6);
    memcpy(pm1, cur, sizeof(double) * 6);
    const double init[6] = { 0, -sw0, 0, 0, sw0, 0 };
    for (int k = 0; k < 6; k++) pj[k] = wrote_next[k] ? nxt[k] : init[k];
    return right_support;
}
