/* TEST INFRASTRUCTURE ONLY (see go1_oracle.h): CPU restatement of the 100 Hz node's glue around PRMPCClass,
 *   xget_position_interpolation()    RT/src/gait_fast.cpp:113-372   (40 Hz -> 100 Hz sample bookkeeping, four cubic calls)
 *   main loop body                   RT/src/gait_fast.cpp:505-746   (counters, swing-foot / foot-rotation calls, the 2 x nh
 *                                    reference windows :568-616, body_theta_mpc :620, the 100-slot /rtMPC/traj :633-729)
 * (RT = unitree_ros/rt_mpc_qp).  The glue is a ROS main() and cannot be compiled; it holds no arithmetic beyond copies,
 * `ref + v * dt`, `/ 5`.  It is restated ONCE here and calls the four class methods through hooks, so the same glue drives
 *   (a) the UNMODIFIED PRMPCClass (oracle/_ref/libref_rt.so: ref_hook_*)  -> tests/golden/rt_node_ref.npz, and
 *   (b) the oracle restatements (orc_hook_* below)                        -> the checker of the device chain.
 *
 * Quirks mirrored on purpose:
 *   - n_t_int = floor(0.025 / 0.01) = 2: the samples shift every 2 fast ticks (50 Hz) although the planner runs at 40 Hz;
 *   - slot 99 of /MPC/Gait carries right_support in {0,1,2} (NLPRTControlClass.cpp:392), so `mpc_gait_flag > 0` gates the
 *     whole fast tick off during left support and `flag > flag_old` fires on support changes, not on every new message;
 *   - rfoot_mpc_ref row 0 is written twice (x, then y) and row 1 never (:585-586, :606-607);
 *   - t_int += floor(count_in_rt_loop / n_t_int) grows quadratically; the interpolation starts once t_int > 2;
 *   - lfoot_inter / rfoot_inter are never updated: _Zsc = 0;
 *   - (int) _height_offset_timex / dt_mpc_fast: the cast binds to the 1, the tick offset is 100.
 * Widened: windows have nh columns (the reference's matrices are 2 x 5 of which body_theta_mpc reads the first _nh = 4).
 */
#include <math.h>
#include <string.h>
#include "go1_oracle.h"

/* node state layout (doubles) for horizon nh: NI = 9 + 3 (nh - 1) */
static int NI_of(int nh) { return 9 + 3 * (nh - 1); }
enum { N_LOOP = 0, N_MPC = 1, N_CNT = 2, N_TINT = 3, N_FLAGOLD = 4, N_COM = 5, N_COMV = 17, N_ACC = 20, N_ZMP = 32, N_DCM = 44, N_INTER = 56 };
int orc_rt_node_doubles(int nh) { return N_INTER + 4 * NI_of(nh) + 6 * (nh + 1) + 6 * nh + 3 + 14; }

void orc_rt_node_default(int nh, double *n)
{
    /* gait_fast.cpp:383-447 */
    memset(n, 0, sizeof(double) * (size_t)orc_rt_node_doubles(nh));
    const double zc = 0.309458, hw = 0.12675;
    for (int q = 0; q < 4; q++) n[N_COM + 3 * q + 2] = zc;              /* COM_in1, COM_in2, COMxyz_ref, COM_ref2: z = Z_C */
    double *inter = n + N_INTER;
    inter[2] = zc;                                                      /* rpy_mpc_body(2) = COM_ref2(2) */
    double *foot = n + N_INTER + 4 * NI_of(nh);                         /* foorpr_gen: y = -+ half hip width for every step */
    for (int j = 0; j < nh + 1; j++) { foot[6 * j + 1] = -hw; foot[6 * j + 4] = hw; }
}

/* one quantity's sample set: in1 | in2 | ref | ref2 (3 each) */
static void shift(double *q) { memcpy(q, q + 3, 3 * sizeof(double)); memcpy(q + 3, q + 6, 3 * sizeof(double)); }

void orc_rt_node_tick(int nh, double *n, const orc_rt_hooks *hk, void *ctx, const double msg[100], int ctrl_flag,
                      const double bodyangle_state[4], double out100[100])
{
    const double dt_fast = 0.01, dt_slow = 0.025;
    const int n_t_int = (int)floor(dt_slow / dt_fast);
    const int NI = NI_of(nh);
    double *COM = n + N_COM, *COMv = n + N_COMV, *ACC = n + N_ACC, *ZMP = n + N_ZMP, *DCM = n + N_DCM;
    double *com_i = n + N_INTER, *acc_i = com_i + NI, *zmp_i = acc_i + NI, *dcm_i = zmp_i + NI;
    double *foot = dcm_i + NI, *rot = foot + 6 * (nh + 1), *thetax = rot + 6 * nh, *bmpc = thetax + 3;
    const double flag = msg[99];
    if (ctrl_flag > 0) {
        n[N_LOOP] += 1;
        n[N_TINT] += (int)floor(n[N_LOOP] / n_t_int);
        if (flag > 0) {
            n[N_MPC] += 1;
            /* ---- xget_position_interpolation ---- */
            n[N_CNT] += 1;
            if (n[N_TINT] > 2) {
                const int w = (int)n[N_CNT];
                hk->mod3(ctx, nh, w, dt_fast, COM, COM + 3, COM + 6, COM + 9, com_i);
                hk->mod3(ctx, nh, w, dt_fast, ACC, ACC + 3, ACC + 6, ACC + 9, acc_i);
                hk->mod3(ctx, nh, w, dt_fast, ZMP, ZMP + 3, ZMP + 6, ZMP + 9, zmp_i);
                hk->mod3(ctx, nh, w, dt_fast, DCM, DCM + 3, DCM + 6, DCM + 9, dcm_i);
            }
            if (((int)n[N_CNT]) % n_t_int == 0) {
                shift(COM); shift(ZMP); shift(DCM); shift(ACC);
                if (flag > n[N_FLAGOLD]) {
                    for (int k = 0; k < 3; k++) {
                        COM[6 + k] = msg[k]; COMv[k] = msg[36 + k];
                        COM[9 + k] = COM[6 + k] + COMv[k] * dt_slow;
                        ACC[6 + k] = msg[39 + k]; ACC[9 + k] = msg[80 + k];
                    }
                    ZMP[6] = msg[12]; ZMP[7] = msg[13]; ZMP[9] = msg[42]; ZMP[10] = msg[43];
                    DCM[6] = msg[34]; DCM[7] = msg[35]; DCM[9] = msg[44]; DCM[10] = msg[45];
                } else {
                    for (int k = 0; k < 3; k++) {
                        COM[6 + k] = msg[k]; COMv[k] = msg[36 + k];
                        COM[6 + k] += COMv[k] * dt_slow;
                        COMv[k] += msg[39 + k] * dt_slow;
                        COM[9 + k] = COM[6 + k] + COMv[k] * dt_slow;
                        ACC[6 + k] = msg[80 + k]; ACC[9 + k] = msg[83 + k];
                    }
                    ZMP[6] = msg[42]; ZMP[7] = msg[43]; ZMP[9] = msg[76]; ZMP[10] = msg[77];
                    DCM[6] = msg[44]; DCM[7] = msg[45]; DCM[9] = msg[78]; DCM[10] = msg[79];
                }
                n[N_CNT] = 0;
                n[N_FLAGOLD] = flag;
            }
            /* ---- swing foot + foot rotation (:534-555) ---- */
            if (n[N_MPC] * dt_fast > 1) {
                const int jf = (int)(n[N_MPC] - (int)1 / dt_fast);
                hk->foot(ctx, nh, jf, 0, msg + 86, foot);
                hk->rot(ctx, nh, jf, dt_fast, rot);
            }
            ZMP[8] = 0;        /* zmpxyz_ref(2) = _Zsc = l/rfoot_inter(2) = 0 */
            /* ---- reference windows (:568-616) ---- */
            double zmp_w[2 * ORC_BODY_NH_MAX], ang_w[2 * ORC_BODY_NH_MAX], rf_w[2 * ORC_BODY_NH_MAX], lf_w[2 * ORC_BODY_NH_MAX], acc_w[ORC_BODY_NH_MAX];
            memset(rf_w, 0, sizeof rf_w);
            for (int jxx = 0; jxx < nh; jxx++) {
                if (jxx == 0) { zmp_w[0] = zmp_i[0]; zmp_w[nh] = zmp_i[1]; acc_w[0] = acc_i[2]; }
                else { zmp_w[jxx] = zmp_i[8 + 3 * jxx - 2]; zmp_w[nh + jxx] = zmp_i[8 + 3 * jxx - 1]; acc_w[jxx] = acc_i[8 + 3 * jxx]; }
                rf_w[jxx] = foot[jxx * 6];
                rf_w[jxx] = foot[jxx * 6 + 1];                  /* row 0 twice, row 1 never */
                lf_w[jxx] = foot[jxx * 6 + 3]; lf_w[nh + jxx] = foot[jxx * 6 + 4];
                ang_w[jxx] = (rot[jxx * 6] + rot[jxx * 6 + 3]) / 5;
                ang_w[nh + jxx] = (rot[jxx * 6 + 1] + rot[jxx * 6 + 4]) / 5;
                if (jxx == 0) { thetax[0] = ang_w[0]; thetax[1] = ang_w[nh]; }
            }
            hk->body(ctx, nh, (int)n[N_MPC], bodyangle_state, zmp_w, ang_w, rf_w, lf_w, acc_w, bmpc);
        }
    }
    /* ---- /rtMPC/traj (:633-729) ---- */
    double inte[51];
    memset(inte, 0, sizeof inte);
    for (int k = 0; k < 3; k++) { inte[k] = com_i[k]; inte[3 + k] = thetax[k]; inte[6 + k] = foot[3 + k]; inte[9 + k] = foot[k]; }
    inte[12] = zmp_i[0]; inte[13] = zmp_i[1]; inte[14] = ZMP[8];
    inte[27] = msg[27];
    for (int k = 0; k < 3; k++) { inte[28 + k] = rot[3 + k]; inte[31 + k] = rot[k]; }
    inte[34] = dcm_i[0]; inte[35] = dcm_i[1];
    for (int k = 0; k < 14; k++) inte[36 + k] = bmpc[k];
    inte[50] = 0;                                                        /* wall time of the tick: not reproduced */
    memset(out100, 0, 100 * sizeof(double));
    for (int k = 0; k < 36; k++) out100[k] = msg[k];
    for (int k = 36; k <= 86; k++) out100[k] = inte[k - 36];
    out100[98] = hk->tx_total(ctx) / 0.001;                              /* (int) binds to _tx_total */
    out100[99] = n[N_LOOP];
}

/* ---------------------------------------------------------------- hooks onto the oracle restatements */
typedef struct {
    int nh;
    double inv[16];
    orc_rt_foot_cfg fc;
    double *foot_state;             /* rt_foot.c layout */
    double *rot_state;              /* bjxx | bjx1 | Rr[3][nh] | Lr[3][nh] */
    orc_body_mpc *body;
} orc_rt_ctx;

static void tables(const orc_rt_ctx *c, double tx[27], double td[27], double *t_end)
{
    const double *ts = c->foot_state;
    tx[0] = 0;
    for (int i = 0; i < 27; i++) td[i] = c->fc.tdsp_ratio * ts[i];
    for (int i = 1; i < 27; i++) { tx[i] = tx[i - 1] + ts[i - 1]; tx[i] = round(tx[i] / c->fc.dt) * c->fc.dt - 0.00001; }
    *t_end = round((tx[26] - 2 * c->fc.tstep) / c->fc.dt_mpc);
}
static void oh_mod3(void *ctx, int nh, int walktime, double dts, const double *a, const double *b, const double *r, const double *r2, double *out)
{
    orc_rt_ctx *c = (orc_rt_ctx *)ctx;
    double tx[27], td[27], t_end;
    tables(c, tx, td, &t_end);      /* the member is Initialize's 3 tstep value before the first swing-foot call: either way >> walktime <= 2 */
    orc_interp_position_mod3(c->inv, nh, (int)t_end, walktime, dts, a, b, r, r2, out);
}
static void oh_foot(void *ctx, int nh, int j, int stop, const double *nrt, double *out)
{
    orc_rt_ctx *c = (orc_rt_ctx *)ctx;
    orc_rt_foot_traj(&c->fc, nh, c->foot_state, j, stop, nrt, out);
}
static void oh_rot(void *ctx, int nh, int j, double dts, double *out)
{
    orc_rt_ctx *c = (orc_rt_ctx *)ctx;
    double tx[27], td[27], t_end;
    tables(c, tx, td, &t_end);
    int bjxx = (int)c->rot_state[0], bjx1 = (int)c->rot_state[1];
    orc_foot_rotation_w(tx, c->foot_state, td, c->foot_state + 27, 0.15, c->fc.dt_mpc, (int)t_end, nh, nh, &bjxx, &bjx1,
                        c->rot_state + 2, c->rot_state + 2 + 3 * nh, j, dts, out);
    c->rot_state[0] = bjxx; c->rot_state[1] = bjx1;
}
static void oh_body(void *ctx, int nh, int i, const double *bs, const double *zmp, const double *ang, const double *rf, const double *lf,
                    const double *acc, double *out14)
{
    orc_rt_ctx *c = (orc_rt_ctx *)ctx;
    (void)nh;
    /* PRMPCClass::body_theta_mpc reads the member _tx, which Foot_trajectory_solve_mod2 rebuilds from the planner's periods */
    double td[27], t_end;
    tables(c, c->body->tx, td, &t_end);
    orc_body_theta_mpc(c->body, i, bs, zmp, ang, rf, lf, acc, out14);
}
static double oh_tx_total(void *ctx)
{
    orc_rt_ctx *c = (orc_rt_ctx *)ctx;
    double tx[27], td[27], t_end;
    tables(c, tx, td, &t_end);
    return (double)(int)tx[26];
}

/* Allocation-free set-up: the caller provides the buffers (foot_state: orc_rt_foot_state_doubles(nh), rot_state: 2 + 6 nh). */
void orc_rt_ctx_init(void *ctx_mem, int nh, double *foot_state, double *rot_state, orc_body_mpc *body)
{
    orc_rt_ctx *c = (orc_rt_ctx *)ctx_mem;
    c->nh = nh;
    orc_interp_aaa_inv_mod(0.025, c->inv);
    orc_rt_foot_cfg_default(&c->fc);
    c->foot_state = foot_state; c->rot_state = rot_state; c->body = body;
    orc_rt_foot_state_default(&c->fc, nh, foot_state);
    memset(rot_state, 0, sizeof(double) * (size_t)(2 + 6 * nh));
    orc_body_cfg bc;
    orc_body_cfg_default(&bc, nh);
    orc_body_init(body, &bc);
}
int orc_rt_ctx_bytes(void) { return (int)sizeof(orc_rt_ctx); }
int orc_body_mpc_bytes(void) { return (int)sizeof(orc_body_mpc); }
void orc_rt_hooks_oracle(orc_rt_hooks *hk) { hk->mod3 = oh_mod3; hk->foot = oh_foot; hk->rot = oh_rot; hk->body = oh_body; hk->tx_total = oh_tx_total; }
