/*
 * grf.c -- oracle restatement of the ground-reaction-force distribution of go1_servo.
 *
 * TEST INFRASTRUCTURE (see go1_oracle.h).  Restates Dynamiccclass (GO1 = unitree_ros/go1_rt_control):
 *   constructor constants / qp_H, qp_h   GO1/src/whole_body_dynamics/dynmics_compute.cpp:55-98
 *   force_distribution                   :141-261  (closed-form split of the planner's L/R wrench)
 *   force_opt                            :265-373  (12-variable QP: 12 equality columns, 24 inequalities)
 *   skew_hat                             :375-385  (frozen bug: vec_w[2,0] is the comma operator, so
 *                                                   every entry uses vec_w[0])
 *   solve_grf_opt                        :387-427
 *   compute_joint_torques                :109-138
 * The equality matrix has identity blocks for the swing legs and all-zero columns for the
 * stance legs: Eigen::QP skips all-zero columns but keeps me = p = 12 (EiQuadProg.cpp:238-241,
 * 288, 370), so inequality constraints that enter the working set can never leave it -- the
 * quirk this QP exercises.  Products keep the shim's evaluation order (ascending inner index).
 */
#include <math.h>
#include <string.h>
#include "go1_oracle.h"

void orc_grf_cfg_default(orc_grf_cfg *c)
{
    c->qp_alpha = 10000; c->qp_beta = 1000; c->qp_gama = 10; c->fz_max = 160; c->mu = 0.25;
}

/* F_leg_ref (3x4 column-major: FR, FL, RR, RL) and F_leg_guess (12) from the planner's wrench
 * F_force_des = (L xyz, R xyz).  :141-261 */
void orc_grf_force_distribution(const double com_des[3], const double leg_des[12], const double F[6], int mode,
                                double yc, const double rfoot_des[3], const double lfoot_des[3],
                                double F_leg_ref[12])
{
#define FR_(r, c) F_leg_ref[(c) * 3 + (r)]
    double dis[4];
    for (int l = 0; l < 4; l++)
        dis[l] = sqrt(pow(com_des[0] - leg_des[3 * l], 2) + pow(com_des[1] - leg_des[3 * l + 1], 2) + pow(com_des[2] - leg_des[3 * l + 2], 2));
    const double FRd = dis[0], FLd = dis[1], RRd = dis[2], RLd = dis[3];
    double f;
    if (mode == 101) {
        f = F[0] * FLd / (FLd + RLd);       FR_(0, 3) = f; FR_(0, 1) = F[0] - f;
        f = F[1] * FLd / (FLd + RLd) * yc;  FR_(1, 3) = f; FR_(1, 1) = F[1] * yc - f;
        f = F[2] * FLd / (FLd + RLd);       FR_(2, 3) = f; FR_(2, 1) = F[2] - f;
        f = F[3] * FRd / (FRd + RRd);       FR_(0, 2) = f; FR_(0, 0) = F[3] - f;
        f = F[4] * FRd / (FRd + RRd) * yc;  FR_(1, 2) = f; FR_(1, 0) = F[4] * yc - f;
        f = F[5] * FRd / (FRd + RRd);       FR_(2, 2) = f; FR_(2, 0) = F[5] - f;
    } else if (mode == 102) {
        double v[3], w[3];
        for (int k = 0; k < 3; k++) { v[k] = leg_des[9 + k] - leg_des[k]; w[k] = lfoot_des[k] - leg_des[k]; }
        double len = sqrt(pow(v[0], 2) + pow(v[1], 2) + pow(v[2], 2));
        double prj = v[0] * w[0] + v[1] * w[1] + v[2] * w[2];
        double r = fmax(fmin(prj / len, 1.0), 0.0);
        f = F[0] * r;       FR_(0, 3) = f; FR_(0, 0) = F[0] - f;
        f = F[1] * r * yc;  FR_(1, 3) = f; FR_(1, 0) = F[1] * yc - f;
        f = F[2] * r;       FR_(2, 3) = f; FR_(2, 0) = F[2] - f;
        for (int k = 0; k < 3; k++) { v[k] = leg_des[6 + k] - leg_des[3 + k]; w[k] = rfoot_des[k] - leg_des[3 + k]; }
        len = sqrt(pow(v[0], 2) + pow(v[1], 2) + pow(v[2], 2));
        prj = v[0] * w[0] + v[1] * w[1] + v[2] * w[2];
        r = fmax(fmin(prj / len, 1.0), 0.0);
        f = F[3] * r;       FR_(0, 2) = f; FR_(0, 1) = F[3] - f;
        f = F[4] * r * yc;  FR_(1, 2) = f; FR_(1, 1) = F[4] * yc - f;
        f = F[5] * r;       FR_(2, 2) = f; FR_(2, 1) = F[5] - f;
    }
#undef FR_
}

/* force_opt + solve_grf_opt (:265-427).  leg_p: FR, FL, RR, RL xyz.  grf: in = previous optimum
 * (enters the gradient and is the solver's untouched x on failure), out = new optimum, or
 * F_leg_guess when the solve produced a NaN.  Returns the solver status. */
int orc_grf_force_opt(const orc_grf_cfg *c, const double base_p[3], const double leg_p[12], const double FT[6],
                      const double F_leg_guess[12], int mode, int right_support, double grf[12],
                      int *active, int *nactive, int *iters, int *qp_solution)
{
    double A[6 * 12];                      /* column-major 6 x 12 */
    memset(A, 0, sizeof A);
    for (int l = 0; l < 4; l++) {
        for (int k = 0; k < 3; k++) A[(3 * l + k) * 6 + k] = 1.0;
        const double w0 = base_p[0] - leg_p[3 * l];        /* skew_hat: every entry reads vec_w[0] */
        double *blk = A + (3 * l) * 6 + 3;                 /* rows 3..5, columns 3l..3l+2 */
        blk[0 * 6 + 0] = 0;    blk[1 * 6 + 0] = -w0;  blk[2 * 6 + 0] = w0;
        blk[0 * 6 + 1] = w0;   blk[1 * 6 + 1] = 0;    blk[2 * 6 + 1] = -w0;
        blk[0 * 6 + 2] = -w0;  blk[1 * 6 + 2] = w0;   blk[2 * 6 + 2] = 0;
    }
    double G[144], Q[144], g0[12];
    /* Q_goal = 2 * (alpha * A' * A + (beta + gama) * I) */
    for (int j = 0; j < 12; j++)
        for (int i = 0; i < 12; i++) {
            double acc = 0.0;
            for (int r = 0; r < 6; r++) acc += (c->qp_alpha * A[i * 6 + r]) * A[j * 6 + r];
            Q[j * 12 + i] = 2 * (acc + (c->qp_beta + c->qp_gama) * (i == j ? 1.0 : 0.0));
        }
    for (int j = 0; j < 12; j++) for (int i = 0; i < 12; i++) G[j * 12 + i] = (Q[i * 12 + j] + Q[j * 12 + i]) / 2.0;
    for (int i = 0; i < 12; i++) {
        double acc = 0.0;
        for (int r = 0; r < 6; r++) acc += (c->qp_alpha * A[i * 6 + r]) * FT[r];
        g0[i] = -2 * ((acc + c->qp_beta * F_leg_guess[i]) + c->qp_gama * grf[i]);
    }
    /* equality columns: identity blocks on the legs whose force must vanish */
    double CE[144], ce0[12];
    memset(CE, 0, sizeof CE); memset(ce0, 0, sizeof ce0);
    int zero_leg[4] = { 0, 0, 0, 0 };     /* FR, FL, RR, RL */
    if (mode == 102) {
        if (right_support == 0) { zero_leg[1] = zero_leg[2] = 1; }
        else if (right_support == 1) { zero_leg[0] = zero_leg[3] = 1; }
    } else if (mode == 101) {
        if (right_support == 0) { zero_leg[0] = zero_leg[2] = 1; }
        else if (right_support == 1) { zero_leg[1] = zero_leg[3] = 1; }
    }
    for (int l = 0; l < 4; l++) if (zero_leg[l]) for (int k = 0; k < 3; k++) CE[(3 * l + k) * 12 + 3 * l + k] = 1.0;
    /* CI = -qp_H', ci0 = qp_h */
    double H[24 * 12], CI[12 * 24], ci0[24];
    memset(H, 0, sizeof H); memset(ci0, 0, sizeof ci0);
#define H_(r, k) H[(r) * 12 + (k)]
    for (int i = 0; i < 4; i++) { H_(2 * i, 3 * i + 2) = -1; H_(2 * i + 1, 3 * i + 2) = 1; ci0[2 * i + 1] = c->fz_max; }
    for (int i = 0; i < 4; i++) { H_(8 + 2 * i, 3 * i) = -1; H_(8 + 2 * i, 3 * i + 2) = -c->mu; H_(8 + 2 * i + 1, 3 * i) = 1; H_(8 + 2 * i + 1, 3 * i + 2) = -c->mu; }
    for (int i = 0; i < 4; i++) { H_(16 + 2 * i, 3 * i + 1) = -1; H_(16 + 2 * i, 3 * i + 2) = -c->mu; H_(16 + 2 * i + 1, 3 * i + 1) = 1; H_(16 + 2 * i + 1, 3 * i + 2) = -c->mu; }
    for (int r = 0; r < 24; r++) for (int k = 0; k < 12; k++) CI[r * 12 + k] = -H_(r, k);
#undef H_
    double X[12], cost;
    memcpy(X, grf, sizeof X);
    int st = orc_qp_solve(12, 12, 24, G, g0, CE, ce0, CI, ci0, X, &cost, active, nactive, iters);
    int ok = 1;
    for (int k = 0; k < 12; k++) if (isnan(X[k])) { ok = 0; break; }
    if (qp_solution) *qp_solution = ok;
    if (ok) memcpy(grf, X, sizeof X); else memcpy(grf, F_leg_guess, sizeof X);
    return st;
}

/* compute_joint_torques :109-138.  Jaco row-major 3x3; swing: PD on the foot, stance: -J' F_ref; + gravity term */
void orc_grf_joint_torques(const double Jaco[9], int swing, const double p_des[3], const double p_est[3],
                           const double pv_des[3], const double pv_est[3], const double F_ref[3],
                           const double grav[3], double swing_kp, double swing_kd, double tau[3])
{
    double w[3];
    for (int k = 0; k < 3; k++) w[k] = swing ? (swing_kp * (p_des[k] - p_est[k]) + swing_kd * (pv_des[k] - pv_est[k])) : F_ref[k];
    for (int i = 0; i < 3; i++) {
        double acc = 0.0;
        for (int r = 0; r < 3; r++) acc += (-Jaco[r * 3 + i]) * w[r];
        tau[i] = acc + grav[i];
    }
}
