/*
 * nlp_node.c -- oracle restatement of the 40 Hz planner node: NLPRTControlClass around NLPClass.
 *
 * TEST INFRASTRUCTURE (see go1_oracle.h).  Restates (NLP = unitree_ros/mosek_nlp_kmp):
 *   NLPRTControlClass::WalkingReactStepping   NLP/src/NLPRTControl/NLPRTControlClass.cpp:191-396  (squat / walk / over
 *                                             branches :196-283, the 100-slot /MPC/Gait layout :288-392)
 *   StartWalking / StopWalking                :400-432
 *   rt_nlp_gait                               :436-596
 *   NLPClass::X_CoM_position_squat            NLP/src/NLP/NLPClass_sqp.cpp:2958-3015, solve_AAA_inv_x :3585-3628
 *   NLPClass::Zmp_distributor                 :3650-3831, zmp_interpolation :3834-3869, Force_torque_calculate :3872-3897
 *   the stop-walking branch of Foot_trajectory_solve_mod2 :2043-2050 (lift heights of the steps ahead zeroed)
 * on top of orc_step_timing_tick_ext and orc_foot_traj_tick (step_timing.c).
 * Pinned bit for bit against the UNMODIFIED NLPRTControlClass (oracle/_ref/libref_nlp.so): tests/golden/rt_node_ref.npz
 * (cfg1: 40 squat ticks, the 671-tick walk, 8 ticks beyond) and tests/golden/nlp_node_ref.npz (stop / restart sequences).
 *
 * Frozen quirks: rt_nlp_gait hands the planner the class member _estimated_state (all zero, never written) instead of its
 * argument, and _feedback_lamda = 0; the foot-feedback arguments do go through.  Force_torque_calculate is fed
 * _thetaaxyx = _Lfootxyzx = _Rfootxyzx = 0 (their setters are commented out in rt_nlp_gait).  A double-support phase whose
 * two ZMP anchors coincide divides 0 / 0: the NaN weights propagate into the force / moment slots.  Slot 98 is never written.
 * The whole-walk arrays _zmpx_real / _zmpy_real become a 64-entry ring (index & 63) with the highest index written so far:
 * an index above it reads the arrays' initial 0.
 */
#include <math.h>
#include <string.h>
#include "go1_oracle.h"

#define NS ORC_FOOTSTEPS

enum { N_ST = 0, N_FOOT = 202, N_RING = 234, N_ZHI = 362, N_LIFT0 = 363, N_RESTART = 364, N_STOP = 365, N_AGAIN = 366,
       N_TINT = 367, N_MPCSTOP = 368, N_RSUP = 369, N_BJX1 = 370, N_PEL = 371, N_LF = 380, N_RF = 383, N_ZMPREF = 386,
       N_DCMREF = 389, N_FL = 392, N_FR = 395, N_ML = 398, N_MR = 401, N_BODY = 404, N_RL = 442, N_COL = 460, N_COR = 463,
       N_COMX = 466, N_COMA = 469, N_TOTAL = 472 };

int orc_nlp_node_doubles(void) { return N_TOTAL; }

void orc_nlp_cfg_default(orc_nlp_cfg *c)
{
    orc_step_cfg_default(&c->step);
    c->dtx = 0.025;                 /* dt_nlp, NLPRTControlClass.h:16 */
    c->height_offset_time = 1.0;    /* :18 */
    c->height_squat_time = 1.0;     /* NLPClass.h:38 */
    c->height_offset = 0.0;         /* NLPClass.h:37 */
    c->z_c = 0.309458;              /* gait::RobotPara_Z_C */
    c->mass = 12.0;
    c->rad = 0.1;                   /* go1: NLPClass_sqp.cpp:262 */
    c->lift_height = 0.03;
    c->steplength = 0.075; c->stepwidth = 2 * c->step.half_hip_width; c->stepheight = 0.0; c->tstep = 0.7;
    c->nsum = 0; c->walkdtime_max = 0;   /* filled by orc_nlp_node_default */
}

void orc_nlp_node_default(orc_nlp_cfg *c, double *n)
{
    memset(n, 0, sizeof(double) * N_TOTAL);
    orc_step_state *st = (orc_step_state *)(n + N_ST);
    orc_step_state_default(st, &c->step, c->steplength, c->stepwidth, c->stepheight, c->tstep);
    orc_foot_state_default(n + N_FOOT, c->stepwidth / 2);       /* _stepwidth(0) = stepwidth / 2, :67 */
    /* _nsum = (_footstepsnumber - 1) _nT, _n_loop_omit = 2 round(_tstep / _dt) (:311);
     * NLPRTControlClass :93: _walkdtime_max = Get_maximal_number(_dtx) + 1 = (_nsum - _n_loop_omit - 1) floor(_dt / _dtx) + 1 */
    c->tx_last0 = st->tx[NS - 1];
    c->nsum = (NS - 1) * (int)round(c->tstep / c->step.dt);          /* NLPClass.h:35-36 */
    const int omit = 2 * (int)round(c->tstep / c->step.dt);
    c->walkdtime_max = (int)((c->nsum - omit - 1) * floor(c->step.dt / c->dtx)) + 1;
    n[N_ZHI] = -1;
    n[N_LIFT0] = NS;                /* first step index whose lift height was zeroed by a stop (none) */
    n[N_RSUP] = 2;
    n[N_PEL + 2] = c->z_c;
    n[N_LF + 1] = c->step.half_hip_width;
    n[N_RF + 1] = -c->step.half_hip_width;
    /* NLPClass::Initialize :560-576 */
    n[N_FR + 2] = 0; n[N_FL + 2] = 0;                       /* the node's own F_R / F_L start at zero (calloc'ed object) */
    n[N_COMX + 2] = c->z_c - c->height_offset;
}

/* NLPRTControlClass::StartWalking / StopWalking :400-432 */
void orc_nlp_node_start(double *n)
{
    if (n[N_STOP] != 0) n[N_AGAIN] = 1;
    n[N_STOP] = 0;
}
void orc_nlp_node_stop(double *n)
{
    if (!(n[N_TINT] < 10)) n[N_STOP] = 1;
}

static void gj7(double *a, double *r)
{
    const int n = 7;
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

/* NLPClass::X_CoM_position_squat :2958-3015: only the vertical entries (2, 5, 8) are ever non-zero */
void orc_nlp_squat(const orc_nlp_cfg *c, int walktime, double dt_sample, double zva[3])
{
    const double t_des = walktime * dt_sample;
    const double tp[3] = { 0.00001, c->height_squat_time / 2 + 0.0001, c->height_squat_time + 0.0001 };
    zva[0] = zva[1] = zva[2] = 0.0;
    if (t_des <= c->height_squat_time) {
        double A[49], Ainv[49];
        const int rowt[7] = { 0, 0, 0, 1, 2, 2, 2 }, kind[7] = { 1, 2, 0, 0, 0, 1, 2 };   /* solve_AAA_inv_x :3585-3628 */
        for (int r = 0; r < 7; r++) {
            const double t = tp[rowt[r]];
            double *a = A + 7 * r;
            if (kind[r] == 0) { a[0] = pow(t, 6); a[1] = pow(t, 5); a[2] = pow(t, 4); a[3] = pow(t, 3); a[4] = pow(t, 2); a[5] = pow(t, 1); a[6] = 1; }
            else if (kind[r] == 1) { a[0] = 6 * pow(t, 5); a[1] = 5 * pow(t, 4); a[2] = 4 * pow(t, 3); a[3] = 3 * pow(t, 2); a[4] = 2 * pow(t, 1); a[5] = 1; a[6] = 0; }
            else { a[0] = 30 * pow(t, 4); a[1] = 20 * pow(t, 3); a[2] = 12 * pow(t, 2); a[3] = 6 * pow(t, 1); a[4] = 2; a[5] = 0; a[6] = 0; }
        }
        gj7(A, Ainv);
        const double plan[7] = { 0, 0, c->z_c, c->z_c - c->height_offset / 2, c->z_c - c->height_offset, 0, 0 };
        double co[7];
        for (int r = 0; r < 7; r++) { double acc = 0.0; for (int k = 0; k < 7; k++) acc += Ainv[7 * r + k] * plan[k]; co[r] = acc; }
        const double t = t_des;
        const double p[7] = { pow(t, 6), pow(t, 5), pow(t, 4), pow(t, 3), pow(t, 2), pow(t, 1), 1 };
        const double v[7] = { 6 * pow(t, 5), 5 * pow(t, 4), 4 * pow(t, 3), 3 * pow(t, 2), 2 * pow(t, 1), 1, 0 };
        const double a[7] = { 30 * pow(t, 4), 20 * pow(t, 3), 12 * pow(t, 2), 6 * pow(t, 1), 2, 0, 0 };
        double z = 0.0, vz = 0.0, az = 0.0;
        for (int k = 0; k < 7; k++) { z += p[k] * co[k]; vz += v[k] * co[k]; az += a[k] * co[k]; }
        zva[0] = z; zva[1] = vz; zva[2] = az;
    } else {
        zva[0] = c->z_c - c->height_offset;
    }
}

static double ring_get(const double *n, int row, int idx)
{
    if (idx < 0 || idx > (int)n[N_ZHI]) return 0.0;          /* never written: the arrays' initial zero */
    return n[N_RING + 64 * row + (idx & 63)];
}

/* NLPClass::Force_torque_calculate :3872-3897 with diagonal _Co_L / _Co_R */
static void force_torque(const orc_nlp_cfg *c, double *n)
{
    const double *com = n + N_COMX, *coma = n + N_COMA, *cl = n + N_COL, *cr = n + N_COR;
    const double j_ini = c->mass * pow(c->rad, 2);
    const double gra[3] = { 0, 0, -c->step.ggg };
    double Ft[3], Lt[3] = { j_ini * 0.0, j_ini * 0.0, j_ini * 0.0 };
    for (int k = 0; k < 3; k++) Ft[k] = c->mass * (coma[k] - gra[k]);
    double *FR = n + N_FR, *FL = n + N_FL, *MR = n + N_MR, *ML = n + N_ML;
    for (int k = 0; k < 3; k++) { FR[k] = cr[k] * Ft[k]; FL[k] = cl[k] * Ft[k]; }
    double rd[3], ld[3];
    for (int k = 0; k < 3; k++) { rd[k] = 0.0 - com[k]; ld[k] = 0.0 - com[k]; }
    const double crR[3] = { FR[1] * rd[2] - FR[2] * rd[1], FR[2] * rd[0] - FR[0] * rd[2], FR[0] * rd[1] - FR[1] * rd[0] };
    const double crL[3] = { FL[1] * ld[2] - FL[2] * ld[1], FL[2] * ld[0] - FL[0] * ld[2], FL[0] * ld[1] - FL[1] * ld[0] };
    for (int k = 0; k < 3; k++) {
        const double Mt = Lt[k] - crR[k] - crL[k];
        MR[k] = cr[k] * Mt; ML[k] = cl[k] * Mt;
    }
}

/* the diagonal of a 3x3 product with a diagonal matrix whose off-diagonal entries are exact zeros: Eigen's row sum adds them */
static void dsp_weights(double *co_a, double *co_b, const double zi[2], const double ze[2], const double zr[2])
{
    double a = fabs(((ze[1] - zi[1]) * (zr[1] - zi[1]) + (ze[0] - zi[0]) * (zr[0] - zi[0])) / (pow(ze[1] - zi[1], 2) + pow(ze[0] - zi[0], 2)));
    double a1 = a;
    if (a > 1) a = 1;
    if (a1 > 1) a1 = 1;
    co_a[0] = a; co_a[1] = a1;
    co_a[2] = sqrt((pow(co_a[0], 2) + pow(co_a[0], 2)) / 2);
    for (int k = 0; k < 3; k++) co_b[k] = 1.0 - co_a[k];
}

/* NLPClass::Zmp_distributor :3650-3831 */
static void zmp_distributor(const orc_nlp_cfg *c, double *n, int walktime, double dt_sample, double zmp_real[2])
{
    const orc_step_state *st = (const orc_step_state *)(n + N_ST);
    const double dt = c->step.dt;
    const int bjx1 = (int)st->bjx1_prev;
    const int j_index = (int)floor(walktime / (dt / dt_sample));
    /* zmp_interpolation :3834-3869 */
    double t_des = walktime * dt_sample - j_index * dt;
    if (t_des <= 0) t_des = 0.0001;
    if (j_index >= 1) {
        zmp_real[0] = (ring_get(n, 0, j_index) - ring_get(n, 0, j_index - 1)) / dt * t_des + ring_get(n, 0, j_index - 1);
        zmp_real[1] = (ring_get(n, 1, j_index) - ring_get(n, 1, j_index - 1)) / dt * t_des + ring_get(n, 1, j_index - 1);
    } else {
        zmp_real[0] = ring_get(n, 0, j_index) / dt * t_des + 0;
        zmp_real[1] = ring_get(n, 1, j_index) / dt * t_des + 0;
    }
    double *col = n + N_COL, *cor = n + N_COR;
    if (bjx1 >= 2) {
        const double tx1 = st->tx[bjx1 - 1], td1 = 0.2 * st->ts[bjx1 - 1];
        const int dsp = (j_index + 1 - round(tx1 / dt)) * dt < td1;
        double *mine = (bjx1 % 2 == 0) ? col : cor, *other = (bjx1 % 2 == 0) ? cor : col;
        if (dsp) {
            const int nTx_n = (int)round(tx1 / dt), nTx_n_dsp = (int)round((tx1 + td1) / dt);
            const double zi[2] = { ring_get(n, 0, nTx_n - 2), ring_get(n, 1, nTx_n - 2) };
            const double ze[2] = { ring_get(n, 0, nTx_n_dsp - 1), ring_get(n, 1, nTx_n_dsp - 1) };
            dsp_weights(mine, other, zi, ze, zmp_real);
        } else {
            for (int k = 0; k < 3; k++) { mine[k] = 1.0; other[k] = 0.0; }
        }
    } else if (bjx1 == 0) {
        for (int k = 0; k < 3; k++) { col[k] = 0.5; cor[k] = 0.5; }
    } else if (bjx1 >= 1) {
        /* _footxyz_real: the step tables with (1, 0) = -stepwidth(0) (:1042-1046) */
        const double zi[2] = { st->footx[bjx1 - 1], (bjx1 - 1 == 0) ? -(c->stepwidth / 2) : st->footy[bjx1 - 1] };
        const double ze[2] = { st->footx[bjx1], st->footy[bjx1] };
        dsp_weights(cor, col, zi, ze, zmp_real);
    }
    force_torque(c, n);
}

/* lift height of step k: NLPClass::FootStepInputs :71-75 and the zeroing of :2043-2048 */
double orc_nlp_lift_ref(double lift_height, int k, int zero_from)
{
    if (k >= zero_from) return 0.0;
    if (k >= NS - 2) return 0.0;
    if (k == NS - 3) return lift_height / 2;
    return lift_height;
}

/* NLPRTControlClass::rt_nlp_gait :436-596 */
static void rt_nlp_gait(const orc_nlp_cfg *c, double *n, int walkdtime1, const double rfoot_fb[3], const double lfoot_fb[3])
{
    orc_step_state *st = (orc_step_state *)(n + N_ST);
    const int t_int = (int)n[N_TINT];
    double *body = n + N_BODY, *rl = n + N_RL;
    if (t_int >= 1) {
        orc_step_in in;
        orc_step_ext ext;
        memset(&in, 0, sizeof in);                       /* _estimated_state: the zero member, see the header */
        memset(&ext, 0, sizeof ext);                     /* _Zsc = 0: flat ground (stepheight 0) */
        in.rfoot_fb[0] = rfoot_fb[0]; in.rfoot_fb[1] = rfoot_fb[1];
        in.lfoot_fb[0] = lfoot_fb[0]; in.lfoot_fb[1] = lfoot_fb[1];
        orc_step_timing_tick_ext(&c->step, t_int, st, &in, body, NULL, &ext);
        for (int q = 0; q < ext.ntdx; q++) {
            n[N_RING + ((t_int + q) & 63)] = ext.zmpx[q];
            n[N_RING + 64 + ((t_int + q) & 63)] = ext.zmpy[q];
        }
        /* entries between the old top and t_int that no tick wrote (a jump ahead) would be stale ring slots: zero them */
        for (int k = (int)n[N_ZHI] + 1; k < t_int; k++) { n[N_RING + (k & 63)] = 0.0; n[N_RING + 64 + (k & 63)] = 0.0; }
        if (t_int + ext.ntdx - 1 > (int)n[N_ZHI]) n[N_ZHI] = t_int + ext.ntdx - 1;
        for (int k = 0; k < 3; k++) { n[N_PEL + k] = body[k]; n[N_PEL + 3 + k] = body[3 + k]; n[N_PEL + 6 + k] = body[6 + k]; }
        n[N_COMX] = body[0]; n[N_COMX + 1] = body[1]; n[N_COMX + 2] = body[2];        /* :1093-1099 */
        n[N_COMA] = body[6]; n[N_COMA + 1] = body[7]; n[N_COMA + 2] = body[8];
        /* Foot_trajectory_solve_mod2 :2043-2048 */
        const int bjx1 = (int)st->bjx1_prev;
        const int t_end = (int)round((c->tx_last0 - 2 * c->tstep) / c->step.dt);
        if (n[N_STOP] != 0 || t_int > t_end) { if (bjx1 + 1 < (int)n[N_LIFT0]) n[N_LIFT0] = bjx1 + 1; }
        const double lift = orc_nlp_lift_ref(c->lift_height, bjx1 - 1, (int)n[N_LIFT0]);
        n[N_RSUP] = orc_foot_traj_tick(&c->step, t_int, st, (int)body[27], n + N_FOOT, c->stepwidth / 2, lift, rl);
        for (int k = 0; k < 3; k++) { n[N_RF + k] = rl[k]; n[N_LF + k] = rl[3 + k]; }
    }
    double zr[2];
    zmp_distributor(c, n, walkdtime1, c->dtx, zr);
    n[N_ZMPREF] = zr[0]; n[N_ZMPREF + 1] = zr[1]; n[N_ZMPREF + 2] = 0.0;             /* zmp_ref = nlp._ZMPxy_realx */
    n[N_ZMPREF] = body[9]; n[N_ZMPREF + 1] = body[10];
    n[N_DCMREF] = body[11]; n[N_DCMREF + 1] = body[12];
    n[N_BJX1] = st->bjx1_prev;
}

/* NLPRTControlClass::WalkingReactStepping :191-396 */
void orc_nlp_node_tick(const orc_nlp_cfg *c, double *n, int walkdtime, int start_mpc, const double rfoot_fb[3],
                       const double lfoot_fb[3], double out100[100])
{
    int walkdtime1 = walkdtime - (int)n[N_RESTART];
    const double hw = c->step.half_hip_width;
    double *pel = n + N_PEL, *lf = n + N_LF, *rf = n + N_RF, *zmp = n + N_ZMPREF, *dcm = n + N_DCMREF, *body = n + N_BODY;
    if (start_mpc) {
        if (n[N_AGAIN] == 0) {
            if (walkdtime1 * c->dtx <= c->height_offset_time) {
                double zva[3];
                orc_nlp_squat(c, walkdtime1, c->dtx, zva);
                for (int k = 0; k < 9; k++) pel[k] = 0.0;
                pel[2] = zva[0]; pel[5] = zva[1]; pel[8] = zva[2];
                for (int k = 0; k < 3; k++) zmp[k] = (lf[k] + rf[k]) / 2;
                for (int k = 0; k < 3; k++) { n[N_FR + k] = 0; n[N_FL + k] = 0; n[N_MR + k] = 0; n[N_ML + k] = 0; }
                n[N_FR + 2] = 9.8 / 2 * c->mass; n[N_FL + 2] = 9.8 / 2 * c->mass;
                lf[0] = 0; lf[1] = hw; lf[2] = 0; rf[0] = 0; rf[1] = -hw; rf[2] = 0;
                for (int k = 0; k < 3; k++) { zmp[k] = (lf[k] + rf[k]) / 2; dcm[k] = zmp[k]; }
                body[13] = zmp[0]; body[14] = zmp[1]; body[15] = dcm[0]; body[16] = dcm[1];
                n[N_BJX1] = 1;
            } else {
                walkdtime1 = (int)(walkdtime1 - (int)c->height_offset_time / c->dtx);
                if (walkdtime1 < c->walkdtime_max) {
                    n[N_TINT] = walkdtime1;
                    rt_nlp_gait(c, n, walkdtime1, rfoot_fb, lfoot_fb);
                } else {
                    n[N_RSUP] = 2;
                    n[N_RESTART] = walkdtime;
                    n[N_MPCSTOP] = 2;
                }
            }
        }
    } else {
        for (int k = 0; k < 3; k++) zmp[k] = (lf[k] + rf[k]) / 2;
        for (int k = 0; k < 3; k++) { n[N_FR + k] = 0; n[N_FL + k] = 0; n[N_MR + k] = 0; n[N_ML + k] = 0; }
        n[N_FR + 2] = 9.8 / 2 * c->mass; n[N_FL + 2] = 9.8 / 2 * c->mass;
        n[N_BJX1] = 1;
    }
    /* :288-392 */
    memset(out100, 0, 100 * sizeof(double));
    for (int k = 0; k < 3; k++) {
        out100[k] = pel[k]; out100[6 + k] = lf[k]; out100[9 + k] = rf[k]; out100[12 + k] = zmp[k];
        out100[15 + k] = n[N_FL + k]; out100[18 + k] = n[N_FR + k]; out100[21 + k] = n[N_ML + k]; out100[24 + k] = n[N_MR + k];
        out100[36 + k] = pel[3 + k]; out100[39 + k] = pel[6 + k];
    }
    out100[27] = n[N_BJX1];
    out100[34] = dcm[0]; out100[35] = dcm[1];
    for (int k = 0; k < 4; k++) out100[42 + k] = body[13 + k];
    for (int k = 0; k < 12; k++) out100[46 + k] = n[N_RL + 6 + k];
    for (int k = 0; k < 21; k++) out100[76 + k] = body[17 + k];
    out100[97] = n[N_MPCSTOP];
    out100[99] = n[N_RSUP];
}
