/*
 * leg_kin.c -- oracle for the Go1 leg kinematics (FK, analytic Jacobian, damped-Newton IK).
 *
 * TEST INFRASTRUCTURE (see go1_oracle.h).  Follows Kinematicclass
 * (GO1 = unitree_ros/go1_rt_control):
 *   constants / leg select   GO1/src/kinematics/Kinematics.cpp:29-41, 65-102
 *   Forward_kinematics       :63-142   (hip frame)
 *   Forward_kinematics_g     :145-229  (world frame, body position + roll/pitch/yaw)
 *   Inverse_kinematics       :233-267  (<= 10 iterations, stop: max(dq) < 1e-4 -- no abs)
 *   Inverse_kinematics_g     :270-304  (<= 15 iterations, stop: |dp|^2 <= 1e-6)
 * The reference writes the world-frame foot position and Jacobian as fully expanded
 * trigonometric polynomials (MATLAB symbolic output, kinematics_matlab/Derive_go1_model.m).
 * This file evaluates the SAME functions in factored form:  p = body_P + Rz(yaw) Ry(pitch)
 * Rx(roll) p_hipframe(q),  J = R J_hipframe(q),  with the planar two-link terms
 * K = l_calf sin(q2+q3) + l_thigh sin q2,  L = l_calf cos(q2+q3) + l_thigh cos q2.
 * It is therefore pinned against the reference's outputs (tests/golden/kin_ref.npz) to
 * 1e-12, not bit for bit; iteration counts of the IK are compared exactly.
 */
#include <math.h>
#include "go1_oracle.h"

static void leg_consts(int leg, double *ox, double *oy, double *ty)
{
    /* 0 FR, 1 FL, 2 RR, 3 RL (any other flag selects RL, as the reference's nested else does) */
    *ox = (leg == 0 || leg == 1) ? 0.1881 : -0.1881;
    *oy = (leg == 0 || leg == 2) ? -0.04675 : 0.04675;
    *ty = (leg == 0 || leg == 2) ? -0.08 : 0.08;
}
#define L_THIGH (-0.213)
#define L_CALF (-0.213)

void orc_leg_fk(const double q[3], int leg, double pos[3], double J[9])
{
    double ox, oy, ty;
    leg_consts(leg, &ox, &oy, &ty);
    const double s1 = sin(q[0]), c1 = cos(q[0]), s2 = sin(q[1]), c2 = cos(q[1]), s3 = sin(q[2]), c3 = cos(q[2]);
    const double s23 = c3 * s2 + c2 * s3, c23 = c3 * c2 - s3 * s2;
    const double K = L_CALF * s23 + L_THIGH * s2;      /* along the hip x axis   */
    const double L = L_CALF * c23 + L_THIGH * c2;      /* along the rotated z    */
    pos[0] = ox + K;
    pos[1] = oy + ty * c1 - s1 * L;
    pos[2] = ty * s1 + c1 * L;
    /* row-major 3x3 */
    J[0] = 0.0;                   J[1] = L;        J[2] = L_CALF * c23;
    J[3] = -(ty * s1 + c1 * L);   J[4] = s1 * K;   J[5] = s1 * (L_CALF * s23);
    J[6] = ty * c1 - s1 * L;      J[7] = -c1 * K;  J[8] = -c1 * (L_CALF * s23);
}

void orc_leg_fk_g(const double bp[3], const double br[3], const double q[3], int leg, double pos[3], double J[9])
{
    double pl[3], Jl[9];
    orc_leg_fk(q, leg, pl, Jl);
    const double sr = sin(br[0]), cr = cos(br[0]), sp = sin(br[1]), cp = cos(br[1]), sy = sin(br[2]), cy = cos(br[2]);
    /* R = Rz(yaw) Ry(pitch) Rx(roll), row-major */
    const double R[9] = { cp * cy, cy * sp * sr - cr * sy, sr * sy + cr * cy * sp,
                          cp * sy, cr * cy + sp * sr * sy, cr * sp * sy - cy * sr,
                          -sp,     cp * sr,                cp * cr };
    for (int i = 0; i < 3; i++) {
        pos[i] = bp[i] + (R[3 * i] * pl[0] + R[3 * i + 1] * pl[1] + R[3 * i + 2] * pl[2]);
        for (int j = 0; j < 3; j++)
            J[3 * i + j] = R[3 * i] * Jl[j] + R[3 * i + 1] * Jl[3 + j] + R[3 * i + 2] * Jl[6 + j];
    }
}

/* dq = lamda J^-1 dp by cofactors (what a fixed-size 3x3 inverse does) */
static void newton_step(const double J[9], const double dp[3], double lamda, double dq[3])
{
    const double c00 = J[4] * J[8] - J[5] * J[7], c01 = J[5] * J[6] - J[3] * J[8], c02 = J[3] * J[7] - J[4] * J[6];
    const double det = J[0] * c00 + J[1] * c01 + J[2] * c02;
    const double id = 1.0 / det;
    const double i00 = c00 * id, i01 = (J[2] * J[7] - J[1] * J[8]) * id, i02 = (J[1] * J[5] - J[2] * J[4]) * id;
    const double i10 = c01 * id, i11 = (J[0] * J[8] - J[2] * J[6]) * id, i12 = (J[2] * J[3] - J[0] * J[5]) * id;
    const double i20 = c02 * id, i21 = (J[1] * J[6] - J[0] * J[7]) * id, i22 = (J[0] * J[4] - J[1] * J[3]) * id;
    dq[0] = (lamda * i00) * dp[0] + (lamda * i01) * dp[1] + (lamda * i02) * dp[2];
    dq[1] = (lamda * i10) * dp[0] + (lamda * i11) * dp[1] + (lamda * i12) * dp[2];
    dq[2] = (lamda * i20) * dp[0] + (lamda * i21) * dp[1] + (lamda * i22) * dp[2];
}

int orc_leg_ik(const double pdes[3], const double qini[3], int leg, double q[3], double J[9])
{
    double pc[3], dp[3], dq[3];
    int it = 0;
    orc_leg_fk(qini, leg, pc, J);
    q[0] = qini[0]; q[1] = qini[1]; q[2] = qini[2];
    for (int j = 0; j < 10; j++) {
        for (int k = 0; k < 3; k++) dp[k] = pdes[k] - pc[k];
        newton_step(J, dp, 0.5, dq);
        const double mx = fmax(dq[0], fmax(dq[1], dq[2]));
        if (mx < 0.0001) break;          /* no abs: frozen quirk, Kinematics.cpp:249 */
        q[0] += dq[0]; q[1] += dq[1]; q[2] += dq[2];
        orc_leg_fk(q, leg, pc, J);
        it++;
    }
    return it;
}

int orc_leg_ik_g(const double bp[3], const double br[3], const double pdes[3], const double qini[3], int leg,
                 double q[3], double J[9])
{
    double pc[3], dp[3], dq[3];
    int it = 0;
    orc_leg_fk_g(bp, br, qini, leg, pc, J);
    q[0] = qini[0]; q[1] = qini[1]; q[2] = qini[2];
    for (int j = 0; j < 15; j++) {
        for (int k = 0; k < 3; k++) dp[k] = pdes[k] - pc[k];
        newton_step(J, dp, 0.5, dq);
        if (fabs(pow(dp[0], 2) + pow(dp[1], 2) + pow(dp[2], 2)) <= 0.000001) break;
        q[0] += dq[0]; q[1] += dq[1]; q[2] += dq[2];
        orc_leg_fk_g(bp, br, q, leg, pc, J);
        it++;
    }
    return it;
}
