/*
 * qp_gi.c -- oracle restatement of the reference's dense dual active-set QP.
 *
 * TEST INFRASTRUCTURE (see go1_oracle.h).  Restates, step for step,
 *   Eigen::QP::solve_quadprog   RT/src/utils/EiQuadProg/EiQuadProg.cpp:493-513
 *   Eigen::QP::solve_quadprog2  RT/src/utils/EiQuadProg/EiQuadProg.cpp:172-491
 *   Eigen::QP::add_constraint   RT/src/utils/EiQuadProg/EiQuadProg.cpp:30-93
 *   Eigen::QP::delete_constraint RT/src/utils/EiQuadProg/EiQuadProg.cpp:95-170
 *   distance / compute_d / update_z / update_r  EiQuadProg.hpp:100-134
 * (RT = unitree_ros/rt_mpc_qp; byte-identical copies live in mosek_nlp_kmp and
 * go1_rt_control.)  The reference's goto flow (l1 / l2 / l2a) is kept as a
 * three-phase loop; every comparison, tie rule (strict '<', first index wins)
 * and tolerance is the reference's.  Deliberately frozen quirks:
 *   - me = p even when all-zero CE columns were skipped (cpp:238-241,288,370)
 *   - equality slots are recorded at A[i] (loop index), not A[iq] (cpp:268)
 *   - 'ss' is NOT reset when the degenerate path jumps back to l2 (cpp:461)
 *   - s[] is not recomputed after the degenerate restore of x (cpp:460-461)
 *   - the stop tolerance uses m (all columns, also never-populated ones)
 *     (cpp:310)
 * Additions (do not change any converged result): an iteration cap on l2a
 * passes (the reference loops forever on a repeating degenerate candidate),
 * zero-initialised A/u (reference reads uninitialised Eigen storage), and
 * status / counters.
 *
 * Eigen primitives are restated in the order documented in DESIGN.md
 * (Appendix "Eigen semantics"): unblocked left-looking lower LLT with true
 * division; row-oriented (dot, then subtract) back-substitution for
 * matrixU().solve; column-oriented substitution for the col-major triangular
 * solves; dot products accumulate in ascending index order; no FMA
 * contraction (build with -ffp-contract=off).
 */
#include <math.h>
#include <string.h>
#include "go1_oracle.h"

#define EPS_D 2.220446049250313e-16

static double gi_hypot(double a, double b)
{
    /* EiQuadProg.hpp:100-118 */
    double a1 = fabs(a), b1 = fabs(b), t;
    if (a1 > b1) { t = b1 / a1; return a1 * sqrt(1.0 + t * t); }
    if (b1 > a1) { t = a1 / b1; return b1 * sqrt(1.0 + t * t); }
    return a1 * sqrt(2.0);
}

typedef struct {
    int n, p, m;
    double *J, *R;                       /* n x n, column-major */
    double *s, *z, *r, *d, *np, *u, *x_old, *u_old;
    int *A, *A_old, *iai, *iaexcl;
} gi_ws;

#define JJ(w, i, j) ((w)->J[(size_t)(j) * (w)->n + (i)])
#define RR(w, i, j) ((w)->R[(size_t)(j) * (w)->n + (i)])

/* d = J' np   (hpp:121-124) */
static void gi_compute_d(gi_ws *w)
{
    int n = w->n;
    for (int j = 0; j < n; j++) {
        double acc = 0.0;
        for (int k = 0; k < n; k++) acc += JJ(w, k, j) * w->np[k];
        w->d[j] = acc;
    }
}

/* z = J[:, iq:] d[iq:]   (hpp:126-129) */
static void gi_update_z(gi_ws *w, int iq)
{
    int n = w->n;
    for (int k = 0; k < n; k++) {
        double acc = 0.0;
        for (int j = iq; j < n; j++) acc += JJ(w, k, j) * w->d[j];
        w->z[k] = acc;
    }
}

/* r[0..iq) = R[0..iq,0..iq)^-1 d[0..iq)   (hpp:131-134), column-oriented */
static void gi_update_r(gi_ws *w, int iq)
{
    for (int i = 0; i < iq; i++) w->r[i] = w->d[i];
    for (int i = iq - 1; i >= 0; i--) {
        w->r[i] = w->r[i] / RR(w, i, i);
        double ri = w->r[i];
        for (int t = 0; t < i; t++) w->r[t] -= ri * RR(w, t, i);
    }
}

static double gi_dot(const double *a, const double *b, int n)
{
    double acc = 0.0;
    for (int i = 0; i < n; i++) acc += a[i] * b[i];
    return acc;
}

/* cpp:30-93 */
static int gi_add_constraint(gi_ws *w, int *iq_io, double *R_norm)
{
    int n = w->n, iq = *iq_io;
    for (int j = n - 1; j >= iq + 1; j--) {
        double cc = w->d[j - 1], ss = w->d[j];
        double h = gi_hypot(cc, ss);
        if (h == 0.0) continue;
        w->d[j] = 0.0;
        ss = ss / h;
        cc = cc / h;
        if (cc < 0.0) { cc = -cc; ss = -ss; w->d[j - 1] = -h; }
        else w->d[j - 1] = h;
        double xny = ss / (1.0 + cc);
        for (int k = 0; k < n; k++) {
            double t1 = JJ(w, k, j - 1), t2 = JJ(w, k, j);
            double a = t1 * cc + t2 * ss;
            JJ(w, k, j - 1) = a;
            JJ(w, k, j) = xny * (t1 + a) - t2;
        }
    }
    iq++;
    for (int i = 0; i < iq; i++) RR(w, i, iq - 1) = w->d[i];
    *iq_io = iq;
    if (fabs(w->d[iq - 1]) <= EPS_D * (*R_norm)) return 0; /* degenerate */
    *R_norm = fmax(*R_norm, fabs(w->d[iq - 1]));
    return 1;
}

/* cpp:95-170; returns 0 if l is not in A[p..iq) (UB in the reference) */
static int gi_delete_constraint(gi_ws *w, int *iq_io, int l)
{
    int n = w->n, p = w->p, iq = *iq_io, qq = -1;
    for (int i = p; i < iq; i++)
        if (w->A[i] == l) { qq = i; break; }
    if (qq < 0) return 0;
    for (int i = qq; i < iq - 1; i++) {
        w->A[i] = w->A[i + 1];
        w->u[i] = w->u[i + 1];
        for (int k = 0; k < n; k++) RR(w, k, i) = RR(w, k, i + 1);
    }
    w->A[iq - 1] = w->A[iq];
    w->u[iq - 1] = w->u[iq];
    w->A[iq] = 0;
    w->u[iq] = 0.0;
    for (int j = 0; j < iq; j++) RR(w, j, iq - 1) = 0.0;
    iq--;
    *iq_io = iq;
    if (iq == 0) return 1;
    for (int j = qq; j < iq; j++) {
        double cc = RR(w, j, j), ss = RR(w, j + 1, j);
        double h = gi_hypot(cc, ss);
        if (h == 0.0) continue;
        cc = cc / h;
        ss = ss / h;
        RR(w, j + 1, j) = 0.0;
        if (cc < 0.0) { RR(w, j, j) = -h; cc = -cc; ss = -ss; }
        else RR(w, j, j) = h;
        double xny = ss / (1.0 + cc);
        for (int k = j + 1; k < iq; k++) {
            double t1 = RR(w, j, k), t2 = RR(w, j + 1, k);
            double a = t1 * cc + t2 * ss;
            RR(w, j, k) = a;
            RR(w, j + 1, k) = xny * (t1 + a) - t2;
        }
        for (int k = 0; k < n; k++) {
            double t1 = JJ(w, k, j), t2 = JJ(w, k, j + 1);
            double a = t1 * cc + t2 * ss;
            JJ(w, k, j) = a;
            JJ(w, k, j + 1) = xny * (a + t1) - t2;
        }
    }
    return 1;
}

/* Eigen LLT<MatrixXd,Lower>::compute, unblocked (n < 32 path): returns 0 on
 * failure.  L is n x n column-major, lower triangle overwritten in place. */
static int gi_llt(double *L, int n)
{
    for (int k = 0; k < n; k++) {
        double x = L[(size_t)k * n + k];
        double sq = 0.0;
        for (int j = 0; j < k; j++) { double v = L[(size_t)j * n + k]; sq += v * v; }
        if (k > 0) x -= sq;
        if (x <= 0.0) return 0;
        x = sqrt(x);
        L[(size_t)k * n + k] = x;
        for (int i = k + 1; i < n; i++) {
            double t = 0.0;
            for (int j = 0; j < k; j++) t += L[(size_t)j * n + i] * L[(size_t)j * n + k];
            double v = L[(size_t)k * n + i];
            if (k > 0) v -= t;
            L[(size_t)k * n + i] = v / x;
        }
    }
    return 1;
}

int orc_qp_solve(int n, int p, int m,
                 const double *G, const double *g0,
                 const double *CE, const double *ce0,
                 const double *CI, const double *ci0,
                 double *x, double *cost,
                 int *active, int *nactive, int *iters)
{
    const double inf = INFINITY;
    const int mp = m + p;
    double Lbuf[n * n], Jbuf[n * n], Rbuf[n * n];
    double sbuf[mp + 1], zbuf[n], rbuf[mp + 1], dbuf[n], npbuf[n], ubuf[mp + 1];
    double xoldbuf[n], uoldbuf[mp + 1], ybuf[n];
    int Abuf[mp + 1], Aoldbuf[mp + 1], iaibuf[mp + 1], iaexclbuf[mp + 1];
    gi_ws W = { n, p, m, Jbuf, Rbuf, sbuf, zbuf, rbuf, dbuf, npbuf, ubuf,
                xoldbuf, uoldbuf, Abuf, Aoldbuf, iaibuf, iaexclbuf };
    gi_ws *w = &W;
    int it_outer = 0, it_add = 0, it_drop = 0, it_degen = 0;
    int status = ORC_OK;

    memset(Abuf, 0, sizeof Abuf); memset(Aoldbuf, 0, sizeof Aoldbuf);
    memset(iaibuf, 0, sizeof iaibuf); memset(iaexclbuf, 0, sizeof iaexclbuf);
    memset(ubuf, 0, sizeof ubuf); memset(uoldbuf, 0, sizeof uoldbuf);
    memset(sbuf, 0, sizeof sbuf); memset(rbuf, 0, sizeof rbuf);
    *nactive = 0;
    iters[0] = iters[1] = iters[2] = iters[3] = 0;

    /* cpp:502 c1 = trace(G) */
    double c1 = 0.0;
    for (int i = 0; i < n; i++) c1 += G[(size_t)i * n + i];

    /* cpp:505-510 */
    memcpy(Lbuf, G, sizeof(double) * n * n);
    if (!gi_llt(Lbuf, n)) { *cost = inf; return ORC_NOT_PD; }
#define LL(i, j) Lbuf[(size_t)(j) * n + (i)]

    /* cpp:207-209 */
    memset(dbuf, 0, sizeof dbuf);
    memset(Rbuf, 0, sizeof Rbuf);
    double R_norm = 1.0;

    /* cpp:213-215  J = L^-T : solve U J = I with U = L' (row-oriented) */
    for (int c = 0; c < n; c++) {
        for (int i = n - 1; i >= 0; i--) {
            double t = 0.0;
            for (int k = i + 1; k < n; k++) t += LL(k, i) * JJ(w, k, c);
            double rhs = (i == c) ? 1.0 : 0.0;
            if (i < n - 1) rhs -= t;
            JJ(w, i, c) = rhs / LL(i, i);
        }
    }
    double c2 = 0.0;
    for (int i = 0; i < n; i++) c2 += JJ(w, i, i);

    /* cpp:227-230  x = -G^-1 g0 */
    for (int i = 0; i < n; i++) ybuf[i] = g0[i];
    for (int i = 0; i < n; i++) {             /* L y = g0, column-oriented */
        ybuf[i] = ybuf[i] / LL(i, i);
        double yi = ybuf[i];
        for (int k = i + 1; k < n; k++) ybuf[k] -= yi * LL(k, i);
    }
    for (int i = n - 1; i >= 0; i--) {        /* L' x = y, row-oriented */
        double t = 0.0;
        for (int k = i + 1; k < n; k++) t += LL(k, i) * ybuf[k];
        double rhs = ybuf[i];
        if (i < n - 1) rhs -= t;
        ybuf[i] = rhs / LL(i, i);
    }
    for (int i = 0; i < n; i++) x[i] = -ybuf[i];
    double f_value = 0.5 * gi_dot(g0, x, n);

    /* cpp:236-276 equality constraints */
    const int me = p, mi = m;
    int iq = 0;
    for (int i = 0; i < me; i++) {
        const double *col = CE + (size_t)i * n;
        int allzero = 1;
        for (int k = 0; k < n; k++) if (!(fabs(col[k]) <= 1e-12)) { allzero = 0; break; }
        if (allzero) continue;
        for (int k = 0; k < n; k++) w->np[k] = col[k];
        gi_compute_d(w);
        gi_update_z(w, iq);
        gi_update_r(w, iq);
        double t2 = 0.0;
        if (fabs(gi_dot(w->z, w->z, n)) > EPS_D)
            t2 = (-gi_dot(w->np, x, n) - ce0[i]) / gi_dot(w->z, w->np, n);
        for (int k = 0; k < n; k++) x[k] += t2 * w->z[k];
        w->u[iq] = t2;
        for (int k = 0; k < iq; k++) w->u[k] -= t2 * w->r[k];
        f_value += 0.5 * (t2 * t2) * gi_dot(w->z, w->np, n);
        w->A[i] = -i - 1;
        if (!gi_add_constraint(w, &iq, &R_norm)) {
            status = ORC_EQ_DEPENDENT;
            goto done;
        }
    }

    for (int i = 0; i < mi; i++) w->iai[i] = i;

    {
        enum { PH_L1, PH_L2, PH_L2A } phase = PH_L1;
        double ss = 0.0, psi, t, t1, t2;
        int ip = 0, l = 0;
        const int cap = 20 * (mp + n) + 50;
        int passes = 0;
        for (;;) {
            if (phase == PH_L1) {
                /* cpp:282-320 */
                it_outer++;
                for (int i = me; i < iq; i++) w->iai[w->A[i]] = -1;
                ss = 0.0; psi = 0.0; ip = 0;
                for (int i = 0; i < mi; i++) {
                    w->iaexcl[i] = 1;
                    double sum = gi_dot(CI + (size_t)i * n, x, n) + ci0[i];
                    w->s[i] = sum;
                    psi += fmin(0.0, sum);
                }
                if (fabs(psi) <= mi * EPS_D * c1 * c2 * 100.0) break;
                for (int i = 0; i < iq; i++) { w->u_old[i] = w->u[i]; w->A_old[i] = w->A[i]; }
                for (int k = 0; k < n; k++) w->x_old[k] = x[k];
                phase = PH_L2;
            }
            if (phase == PH_L2) {
                /* cpp:322-342 */
                for (int i = 0; i < mi; i++)
                    if (w->s[i] < ss && w->iai[i] != -1 && w->iaexcl[i]) { ss = w->s[i]; ip = i; }
                if (ss >= 0.0) break;
                for (int k = 0; k < n; k++) w->np[k] = CI[(size_t)ip * n + k];
                w->u[iq] = 0.0;
                w->A[iq] = ip;
                phase = PH_L2A;
            }
            /* PH_L2A: cpp:349-490 */
            if (++passes > cap) { status = ORC_ITER_CAP; break; }
            gi_compute_d(w);
            gi_update_z(w, iq);
            gi_update_r(w, iq);
            l = 0;
            t1 = inf;
            for (int k = me; k < iq; k++) {
                double tmp;
                if (w->r[k] > 0.0 && ((tmp = w->u[k] / w->r[k]) < t1)) { t1 = tmp; l = w->A[k]; }
            }
            if (fabs(gi_dot(w->z, w->z, n)) > EPS_D)
                t2 = -w->s[ip] / gi_dot(w->z, w->np, n);
            else
                t2 = inf;
            t = fmin(t1, t2);
            if (t >= inf) { status = ORC_INFEASIBLE; f_value = inf; break; }   /* case (i) */
            if (t2 >= inf) {                                                   /* case (ii) */
                for (int k = 0; k < iq; k++) w->u[k] -= t * w->r[k];
                w->u[iq] += t;
                w->iai[l] = l;
                if (!gi_delete_constraint(w, &iq, l)) { status = ORC_ITER_CAP; break; }
                it_drop++;
                continue; /* l2a */
            }
            /* case (iii) */
            {
                double zn = gi_dot(w->z, w->np, n);
                for (int k = 0; k < n; k++) x[k] += t * w->z[k];
                f_value += t * zn * (0.5 * t + w->u[iq]);
            }
            for (int k = 0; k < iq; k++) w->u[k] -= t * w->r[k];
            w->u[iq] += t;
            if (t == t2) {
                if (!gi_add_constraint(w, &iq, &R_norm)) {
                    /* cpp:444-462 degenerate */
                    it_degen++;
                    w->iaexcl[ip] = 0;
                    if (!gi_delete_constraint(w, &iq, ip)) { status = ORC_ITER_CAP; break; }
                    for (int i = 0; i < m; i++) w->iai[i] = i;
                    for (int i = 0; i < iq; i++) {
                        w->A[i] = w->A_old[i];
                        w->iai[w->A[i]] = -1;
                        w->u[i] = w->u_old[i];
                    }
                    for (int k = 0; k < n; k++) x[k] = w->x_old[k];
                    phase = PH_L2;
                    continue;
                }
                it_add++;
                w->iai[ip] = -1;
                phase = PH_L1;
                continue;
            }
            /* partial step: cpp:477-490 */
            w->iai[l] = l;
            if (!gi_delete_constraint(w, &iq, l)) { status = ORC_ITER_CAP; break; }
            it_drop++;
            w->s[ip] = gi_dot(CI + (size_t)ip * n, x, n) + ci0[ip];
            /* stay in PH_L2A */
        }
    }

done:
    *cost = f_value;
    *nactive = iq;
    for (int i = 0; i < iq && i < mp; i++) active[i] = w->A[i];
    iters[ORC_IT_OUTER] = it_outer; iters[ORC_IT_ADD] = it_add;
    iters[ORC_IT_DROP] = it_drop; iters[ORC_IT_DEGEN] = it_degen;
    if (status == ORC_OK || status == ORC_EQ_DEPENDENT)
        for (int i = 0; i < n; i++) if (isnan(x[i])) { status = ORC_NAN; break; }
    return status;
#undef LL
}
