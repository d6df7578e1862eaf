/* TEST INFRASTRUCTURE ONLY (see go1_oracle.h): CPU restatement of the swing-foot roll / pitch reference of rt_mpc_qp
 * (SURVEY.md 8f row 2), pinned against the unmodified class by tests/test_oracle_vs_ref.py (live oracle/_ref + golden).
 *
 *   PRMPCClass::XGetSolution_Foot_rotation  RT/src/FastMPC/PRMPCClass.cpp:2255-2380 (called at gait_fast.cpp:544)
 *   PRMPCClass::Indexfind (first branch)    :716-727, clamped to the 27-entry table as in body_mpc.c
 *
 * Per horizon sample jx (tick walktimex + jx): step indices from the step-time table, then for the swing foot of that
 * step a roll bump  -0.065 / +0.075 (1 - cos(2 pi s / T))  and, in the second half of the step and only for a forward
 * step, a pitch  0.075 dx / footx_max (cos(4 pi s / T) - 1);  first half: pitch 0.
 * Frozen quirks: the angle arrays are MEMBERS and a branch that does not write leaves the value an earlier call put
 * into that column (backward or zero-length steps keep the last pitch; the stance foot keeps its last swing values);
 * yaw is column 0 of row 2 for every sample (never written: 0); beyond _t_end_footstep the step index of the last
 * in-range tick stays and nothing is written. */
#include <math.h>
#include <string.h>
#include "go1_oracle.h"

static int rot_indexfind(const double *tx, double goal)
{
    int j = 0;
    while (j < 27 && goal >= tx[j]) j++;
    return j - 1;
}

void orc_foot_rot_state_init(orc_foot_rot_state *s) { memset(s, 0, sizeof *s); }

/* General form: the angle members are 3 x ncol row-major (the reference's are 3 x 5), nh <= ncol samples. */
void orc_foot_rotation_w(const double tx[27], const double ts[27], const double td[27], const double footx[27], double footx_max,
                         double dt_mpc, int t_end_footstep, int nh, int ncol, int *bjxx_io, int *bjx1_io, double *Rr, double *Lr,
                         int walktimex, double dt_sample, double *out)
{
    int bjxx = *bjxx_io, bjx1 = *bjx1_io;
    memset(out, 0, sizeof(double) * 6 * (size_t)nh);
    for (int walktime = walktimex; walktime < walktimex + nh; walktime++) {
        const int col = walktime - walktimex;
        if (walktime <= t_end_footstep) {
            bjxx = rot_indexfind(tx, walktime * dt_mpc) + 1;
            bjx1 = rot_indexfind(tx, (walktime + 1) * dt_mpc) + 1;
        }
        const int k = bjx1 - 1;
        if (bjx1 >= 2 && walktime <= t_end_footstep) {
            const double t_des = (walktime + 1) * dt_sample - (tx[k] + 2 * td[k] / 4);
            const double sarg = t_des + 2 * td[k] / 4;
            const double dfx = footx[bjx1] - footx[bjx1 - 1];
            double *A = (bjx1 % 2 == 0) ? Rr : Lr;          /* even: right foot swings */
            const double amp = (bjx1 % 2 == 0) ? -0.065 : 0.075;
            A[0 * ncol + col] = amp * (1 - cos(2 * M_PI / ts[k] * sarg));
            if (sarg >= ts[k] / 2) {
                if (dfx > 0) A[1 * ncol + col] = 0.075 * dfx / footx_max * (cos(4 * M_PI / ts[k] * sarg) - 1);
            } else {
                A[1 * ncol + col] = 0;
            }
        }
        out[6 * col + 0] = Rr[0 * ncol + col]; out[6 * col + 1] = Rr[1 * ncol + col]; out[6 * col + 2] = Rr[2 * ncol + 0];
        out[6 * col + 3] = Lr[0 * ncol + col]; out[6 * col + 4] = Lr[1 * ncol + col]; out[6 * col + 5] = Lr[2 * ncol + 0];
    }
    *bjxx_io = bjxx; *bjx1_io = bjx1;
}

/* out[6 * nh]: per sample  right roll, pitch, yaw | left roll, pitch, yaw.  nh <= 5 (the reference's 3x5 members). */
void orc_foot_rotation(const double tx[27], const double ts[27], const double td[27], const double footx[27], double footx_max,
                       double dt_mpc, int t_end_footstep, int nh, orc_foot_rot_state *s, int walktimex, double dt_sample, double *out)
{
    orc_foot_rotation_w(tx, ts, td, footx, footx_max, dt_mpc, t_end_footstep, nh, 5, &s->bjxx, &s->bjx1, s->Rr, s->Lr, walktimex, dt_sample, out);
}
