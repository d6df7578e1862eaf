/* TEST INFRASTRUCTURE ONLY (see go1_oracle.h): CPU restatement of the 40 Hz -> 100 Hz reference
 * interpolation of rt_mpc_qp, pinned against the unmodified class by
 * tests/test_oracle_vs_ref.py::test_oracle_ref_interp_vs_reference (live oracle/_ref + golden vectors).
 *
 *   PRMPCClass::solve_AAA_inv_mod1         RT/src/FastMPC/PRMPCClass.cpp:1344-1361
 *   PRMPCClass::XGetSolution_position_mod3 RT/src/FastMPC/PRMPCClass.cpp:1170-1261
 *
 * gait_fast.cpp:131-138 calls it once per interpolated quantity (CoM, CoM acceleration, ZMP, DCM): a cubic
 * through the samples at t = -dt, 0, dt, 2dt, evaluated at the horizon's nh instants.  The result holds
 * position / velocity / acceleration at the first instant and the positions at the nh - 1 later ones --
 * the reference rows the body-inclination MPC takes.
 *
 * Frozen as the reference has them: libm pow for the monomials (also pow(t, 1), pow(t, 0)), the inverse by
 * row-pivoted Gauss-Jordan (what oracle/eigen_shim does above 3x3; real Eigen's 4x4 cofactor form may
 * differ in the last bits -- DESIGN.md section 4), products left to right with ascending inner index. */
#include <math.h>
#include <string.h>
#include "go1_oracle.h"

/* _AAA_inv_mod: inverse of the Vandermonde rows (t^3, t^2, t, 1) at t = -dt, 0, dt, 2dt.  Row-major 4x4. */
void orc_interp_aaa_inv_mod(double dt, double inv[16])
{
    const double t[4] = {-dt, 0.0, dt, 2 * dt};
    double a[4][4], r[4][4];
    for (int i = 0; i < 4; i++) {
        a[i][0] = pow(t[i], 3); a[i][1] = pow(t[i], 2); a[i][2] = pow(t[i], 1); a[i][3] = 1.0;
        for (int j = 0; j < 4; j++) r[i][j] = (i == j) ? 1.0 : 0.0;
    }
    for (int k = 0; k < 4; k++) {
        int piv = k;
        double best = fabs(a[k][k]);
        for (int i = k + 1; i < 4; i++) if (fabs(a[i][k]) > best) { best = fabs(a[i][k]); piv = i; }
        if (piv != k)
            for (int j = 0; j < 4; j++) {
                double s = a[k][j]; a[k][j] = a[piv][j]; a[piv][j] = s;
                s = r[k][j]; r[k][j] = r[piv][j]; r[piv][j] = s;
            }
        const double d = a[k][k];
        for (int j = 0; j < 4; j++) { a[k][j] = a[k][j] / d; r[k][j] = r[k][j] / d; }
        for (int i = 0; i < 4; i++) {
            if (i == k) continue;
            const double f = a[i][k];
            for (int j = 0; j < 4; j++) { a[i][j] -= f * a[k][j]; r[i][j] -= f * r[k][j]; }
        }
    }
    memcpy(inv, r, sizeof r);
}

static double row_inv_temp(const double row[4], const double inv[16], const double temp[4])
{
    double v[4];
    for (int j = 0; j < 4; j++) {          /* (1x4 * 4x4) first ... */
        double acc = 0.0;
        for (int k = 0; k < 4; k++) acc += row[k] * inv[4 * k + j];
        v[j] = acc;
    }
    double acc = 0.0;                      /* ... then * 4x1 */
    for (int k = 0; k < 4; k++) acc += v[k] * temp[k];
    return acc;
}

/* out[9 + 3 (nh - 1)]: [0,3) position, [3,6) velocity, [6,9) acceleration at walktime * dt_sample; then the xyz
 * positions at the following nh - 1 samples.  All zero once walktime > t_end_footstep. */
void orc_interp_position_mod3(const double inv[16], int nh, int t_end_footstep, int walktime, double dt_sample,
                              const double in1[3], const double in2[3], const double ref[3], const double ref2[3],
                              double *out)
{
    memset(out, 0, sizeof(double) * (size_t)(9 + 3 * (nh - 1)));
    if (!(walktime <= t_end_footstep)) return;
    for (int jx = 0; jx < nh; jx++) {
        const double t = walktime * dt_sample + jx * dt_sample;
        const double p[4] = {pow(t, 3), pow(t, 2), pow(t, 1), pow(t, 0)};
        const double v[4] = {3 * pow(t, 2), 2 * pow(t, 1), 1, 0};
        const double a[4] = {6 * pow(t, 1), 2, 0, 0};
        for (int c = 0; c < 3; c++) {
            const double temp[4] = {in1[c], in2[c], ref[c], ref2[c]};
            if (jx == 0) {
                out[c] = row_inv_temp(p, inv, temp);
                out[3 + c] = row_inv_temp(v, inv, temp);
                out[6 + c] = row_inv_temp(a, inv, temp);
            } else {
                out[8 + 3 * jx - 2 + c] = row_inv_temp(p, inv, temp);
            }
        }
    }
}
