/*
 * go1_oracle.h -- CPU oracle for the Go1 gait-planning MPC hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a dependency-free C restatement of the
 * reference's CPU algorithm (jtdingx/quadrupedal_loco), used as the checker by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs.  Nothing under quadrupedal_loco_b200/ may include, link or
 * call it: the product path is CUDA-only and fails loudly without its
 * extension.
 *
 * Parity status: the reference ships no tests or golden vectors for this path
 * (SURVEY.md section 4), and Eigen 3 -- the un-vendored, version-unpinned
 * system package whose LLT / triangular solves / dot products the reference
 * calls -- is absent from this image.  The pin that exists is
 * oracle/_ref: the UNMODIFIED reference sources (EiQuadProg.cpp,
 * QPBaseClass.cpp, Kinematics.cpp) compiled against oracle/eigen_shim (a
 * minimal stand-in for the Eigen headers written for this repo) and compared
 * bit-for-bit with this restatement in tests/test_oracle_vs_ref.py (run in the
 * authoring container, golden outputs committed under tests/golden/).  That
 * pins control flow, tie-breaking, tolerances and every reference quirk; it
 * cannot pin Eigen's own floating-point summation order, which is stated in
 * DESIGN.md ("parity pinned up to Eigen's summation order").
 *
 * Conventions (same as the reference, RT/src/utils/EiQuadProg/EiQuadProg.hpp:15-32):
 *     min 0.5 x'Gx + g0'x   s.t.  CE'x + ce0 = 0,  CI'x + ci0 >= 0
 * All matrices column-major double; CE is n x p, CI is n x m (one constraint
 * per column).
 */
#ifndef GO1_ORACLE_H
#define GO1_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* status codes shared with include/go1mpc.h (kept numerically identical) */
enum {
    ORC_OK = 0,            /* converged, cost finite                         */
    ORC_NOT_PD = 1,        /* LLT failed: x untouched, cost = +inf           */
    ORC_INFEASIBLE = 2,    /* t = +inf in step 2c: cost = +inf, x is garbage */
    ORC_ITER_CAP = 3,      /* safety cap hit (reference has no cap)          */
    ORC_NAN = 4,           /* NaN in x after the solve (solveQP() == false)  */
    ORC_EQ_DEPENDENT = 5   /* equality add failed; early return as reference */
};

/* iteration counters written to iters[4] */
enum { ORC_IT_OUTER = 0, ORC_IT_ADD = 1, ORC_IT_DROP = 2, ORC_IT_DEGEN = 3 };

/*
 * Goldfarb-Idnani dual active-set solve.
 * Follows Eigen::QP::solve_quadprog / solve_quadprog2 / add_constraint /
 * delete_constraint, RT/src/utils/EiQuadProg/EiQuadProg.cpp:30-513.
 *   x        in/out, n (left untouched when G is not PD, like the reference)
 *   cost     out, objective value or +inf
 *   active   out, m+p ints: final working set A[0..nactive) (equalities are
 *            stored as -i-1, inequalities by column index)
 *   iters    out, 4 counters (ORC_IT_*)
 * returns a status code (ORC_*).
 */
int orc_qp_solve(int n, int p, int m,
                 const double *G, const double *g0,
                 const double *CE, const double *ce0,
                 const double *CI, const double *ci0,
                 double *x, double *cost,
                 int *active, int *nactive, int *iters);

/* ------------------------------------------------------------------------
 * Body-inclination MPC (PRMPCClass), horizon-parametrised.
 * Follows RT/src/FastMPC/PRMPCClass.cpp:46-374 (model part), 379-849.
 * --------------------------------------------------------------------- */
#define ORC_FOOTSTEPS 27
#define ORC_BODY_NH_MAX 40

typedef struct {
    int nh;                 /* horizon (reference: 4)                        */
    double dt_mpc;          /* 0.01                                          */
    double dt_slow;         /* 0.025 (used to round _tx)                     */
    double tstep;           /* 0.7                                           */
    double height_offset_time; /* 1.0                                        */
    double g, mass, j_ini;  /* 9.8, 12, 0.12                                 */
    double foot_length, foot_width; /* 0.02, 0.02 (PRMPC members)            */
    double theta_lim;       /* 10 deg                                        */
    double torque_lim;      /* 20 (divided by j_ini as in the reference)     */
    double Rtheta, alphatheta, beltatheta, gama_zmp; /* 100, 10, 5e9, 5000   */
    double lamda[4];        /* feedback gains lamdax, lamdavx, lamday, lamdavy (0) */
} orc_body_cfg;

typedef struct {
    orc_body_cfg cfg;
    /* model matrices (column-major nh x nh / nh x 2) */
    double pps[ORC_BODY_NH_MAX * 2], pvs[ORC_BODY_NH_MAX * 2];
    double ppu[ORC_BODY_NH_MAX * ORC_BODY_NH_MAX], pvu[ORC_BODY_NH_MAX * ORC_BODY_NH_MAX];
    double ppu_2[ORC_BODY_NH_MAX * ORC_BODY_NH_MAX], pvu_2[ORC_BODY_NH_MAX * ORC_BODY_NH_MAX];
    int nstepx, nsum_mpc;
    /* per-instance state */
    double tx[ORC_FOOTSTEPS];
    double thetaxk[2], thetayk[2];
    double V_ini[2 * ORC_BODY_NH_MAX];
    /* rolled-out members that survive between ticks (gated ticks return them) */
    double thetax[ORC_BODY_NH_MAX], thetay[ORC_BODY_NH_MAX];
    double zmpx_real[ORC_BODY_NH_MAX], zmpy_real[ORC_BODY_NH_MAX];
    double torquex_real0, torquey_real0;
    int bjx1, bjx2;
    int qp_solution;
    /* diagnostics of the last solve */
    int status, nactive, iters[4];
    int active[12 * ORC_BODY_NH_MAX];
    double cost;
} orc_body_mpc;

void orc_body_cfg_default(orc_body_cfg *c, int nh);
void orc_body_init(orc_body_mpc *s, const orc_body_cfg *c);
/* refs are row-major-by-signal: zmp_ref[2*nh] = row0 (x) then row1 (y), etc.
 * comacc_z_ref[nh] is row 2 of the reference's 3 x nh comacc matrix.        */
void orc_body_theta_mpc(orc_body_mpc *s, int i, const double bodyangle_state[4],
                        const double *zmp_ref, const double *bodyangle_ref,
                        const double *rfoot_ref, const double *lfoot_ref,
                        const double *comacc_z_ref, double out14[14]);

/* flat batch helper used by the CPU baseline: B independent instances, each
 * with its own state; single thread. Layouts documented in body_mpc.c.      */
void orc_body_step_batch(const orc_body_cfg *c, int B, const int *tick,
                         const double *tx, double *theta_state /*B*4*/,
                         const double *bodyangle_state /*B*4*/,
                         const double *refs /*B*9*nh*/, double *out14 /*B*14*/,
                         double *x_out /*B*2nh*/, int *active /*B*12nh*/,
                         int *nactive, int *iters /*B*4*/, int *status);


/* ------------------------------------------------------------------------
 * Step-location / step-timing SQP (NLPClass), one 40 Hz tick.
 * Follows NLP/src/NLP/NLPClass_sqp.cpp (NLP = unitree_ros/mosek_nlp_kmp):
 *   step_timing_opti_loop :693-1102 (SQP loop :776-810, write-back :886-916,
 *   LIPM roll-out :938-955, feedback blend :1017-1022, indices :1031-1041),
 *   Indexfind :1105-1141, step_timing_object_function :1144-1173,
 *   step_timing_constraints :1175-1458, solve_stepping_timing :1613-1639.
 *   CoM_height_solve :2361-2473 (6th-order vertical CoM polynomial, 7x7 inverse).
 * With cfg.ext_height != 0 the vertical CoM samples are taken from the inputs
 * instead (comz/comaz for ticks i..i+2, comvz at i).
 * --------------------------------------------------------------------- */
#define ORC_STEP_NQP_MAX 8
typedef struct {
    double dt, Wn, ggg;             /* 0.025, sqrt(g/hcom), 9.8                   */
    double t_min, t_max;            /* 0.5, 1.0                                   */
    double footx_max, footx_min;    /* 0.15, -0.05                                */
    double footx_vmax, footx_vmin, footy_vmax, footy_vmin;   /* 3, -2.875, 2, -1  */
    double comax_max, comax_min, comay_max, comay_min;       /* 5, -5, 6, -6      */
    double aax, aay, aaxv, aayv, bbx, bby, rr1, rr2;         /* go1 weights       */
    double half_hip_width, foot_width;                        /* 0.12675, 0.03     */
    double lamda[4];                /* comx, comvx, comy, comvy feedback gains (0) */
    double hcom;                    /* 0.309458 - 0 (RobotPara_Z_C - _height_offset) */
    int n_sqp;                      /* 3                                          */
    int ext_height;                 /* 0: CoM_height_solve; 1: comz/comaz/comvz0 from the inputs */
} orc_step_cfg;

typedef struct {                    /* carried from tick to tick (202 doubles)    */
    double ts[ORC_FOOTSTEPS], tx[ORC_FOOTSTEPS];
    double footx[ORC_FOOTSTEPS], footy[ORC_FOOTSTEPS], footz[ORC_FOOTSTEPS];
    double Lxx[ORC_FOOTSTEPS], Lyy[ORC_FOOTSTEPS];
    double feed[6];                 /* com x, vx, ax, y, vy, ay _feed at tick i-1  */
    double vari[4];                 /* _Vari_ini.col(i-1) = [Lx, Ly, tr1, tr2]     */
    double endref[2];               /* _comvx_endref, _comvy_endref                */
    double bjx1_prev;               /* _bjx1 left by the previous tick (exact integer) */
} orc_step_state;

typedef struct {                    /* per-tick inputs (20 doubles)               */
    double est[6];                  /* estimated com x, vx, ax, y, vy, ay          */
    double rfoot_fb[2], lfoot_fb[2];/* measured foot x, y                          */
    double comz[3], comaz[3], zsc[3], comvz0;
} orc_step_in;

typedef struct {
    int periond_i, k_yu, bjxx, bjx1, n_solved;
    int status[ORC_STEP_NQP_MAX], nactive[ORC_STEP_NQP_MAX], iters[ORC_STEP_NQP_MAX][4];
    int active[ORC_STEP_NQP_MAX][25];
    double x[ORC_STEP_NQP_MAX][4];  /* the QP increment of each SQP iteration      */
} orc_step_diag;

/* the part of the LIPM roll-out (:938-955) beyond sample i + 2: the reference rolls out jxx = 1.._nTdx samples per tick, and
 * NLPClass::Zmp_distributor reads the ZMP of the later ones.  zsc: terrain height _Zsc(i + q), input (q >= 3 used). */
#define ORC_NTD_MAX 12
typedef struct { int ntdx; double zsc[ORC_NTD_MAX]; double zmpx[ORC_NTD_MAX], zmpy[ORC_NTD_MAX]; } orc_step_ext;

void orc_step_cfg_default(orc_step_cfg *c);
/* default tables of NLPClass::FootStepInputs/Initialize (:51-75,:160-206) */
void orc_step_state_default(orc_step_state *s, const orc_step_cfg *c, double steplength, double stepwidth,
                            double stepheight, double tstep);
void orc_step_timing_tick(const orc_step_cfg *c, int i, orc_step_state *s, const orc_step_in *in,
                          double out38[38], orc_step_diag *diag);
void orc_step_timing_tick_ext(const orc_step_cfg *c, int i, orc_step_state *s, const orc_step_in *in,
                              double out38[38], orc_step_diag *diag, orc_step_ext *ext);
/* flat batch driver: states [B][202], ins [B][20], out [B][38], diag optional */
void orc_step_timing_batch(const orc_step_cfg *c, int B, const int *tick, double *states, const double *ins,
                           double *out38, orc_step_diag *diag);

/* ------------------------------------------------------------------------
 * Swing-foot trajectory of the step planner: NLPClass::Foot_trajectory_solve_mod2
 * (NLP/src/NLP/NLPClass_sqp.cpp:2039-2358) with solve_AAA_inv2 (:3633-3645), called
 * right after step_timing_opti_loop with the same tick (NLPRTControlClass.cpp:470).
 * The stop-walking branch (:2043-2048) is the caller's: it passes the lift height of the current step (nlp_node.c).
 * The reference keeps whole-walk arrays _R/Lfoot{x,y,z}; only a sliding window
 * is ever read, carried here as 32 doubles:
 *   [0,6)  R xyz, L xyz at tick j-1      [6,12)  values the arrays hold at j before this tick
 *   [12,18) at j-2   [18,24) at j-3      [24,30) value frozen at the step's start (index s-2)
 *   [30] s the freeze belongs to (-1: none)   [31] _ry_left_right
 * --------------------------------------------------------------------- */
#define ORC_FOOT_STATE 32
void orc_foot_state_default(double fs[ORC_FOOT_STATE], double stepwidth0);
/* st: the planner state AFTER this tick's orc_step_timing_tick; bjxx: its out38[27].
 * out18 = the Vec18 of the reference; returns right_support (0, 1 or 2). */
int orc_foot_traj_tick(const orc_step_cfg *c, int j, const orc_step_state *st, int bjxx,
                       double fs[ORC_FOOT_STATE], double stepwidth0, double lift_height, double out18[18]);

/* ------------------------------------------------------------------------
 * The 40 Hz planner node (nlp_node.c): NLPRTControlClass::WalkingReactStepping / rt_nlp_gait / StartWalking / StopWalking
 * (NLP/src/NLPRTControl/NLPRTControlClass.cpp:191-596) with NLPClass::X_CoM_position_squat (NLPClass_sqp.cpp:2958-3015),
 * Zmp_distributor / zmp_interpolation / Force_torque_calculate (:3650-3897).  The node is a flat array of
 * orc_nlp_node_doubles() doubles: [0,202) planner state | [202,234) swing-foot window | ZMP ring | members.
 * --------------------------------------------------------------------- */
typedef struct {
    orc_step_cfg step;
    double dtx, height_offset_time, height_squat_time, height_offset, z_c, mass, rad, lift_height;
    double steplength, stepwidth, stepheight, tstep;
    double tx_last0;                /* _tx(last) as Initialize left it: _t_end_footstep = round((tx_last0 - 2 tstep) / dt) */
    int nsum, walkdtime_max;
} orc_nlp_cfg;
void orc_nlp_cfg_default(orc_nlp_cfg *c);
int orc_nlp_node_doubles(void);
void orc_nlp_node_default(orc_nlp_cfg *c, double *node);        /* also fills c->nsum, walkdtime_max, tx_last0 */
void orc_nlp_node_start(double *node);
void orc_nlp_node_stop(double *node);
void orc_nlp_squat(const orc_nlp_cfg *c, int walktime, double dt_sample, double zva[3]);
double orc_nlp_lift_ref(double lift_height, int k, int zero_from);
void orc_nlp_node_tick(const orc_nlp_cfg *c, double *node, int walkdtime, int start_mpc, const double rfoot_fb[3],
                       const double lfoot_fb[3], double out100[100]);

/* ------------------------------------------------------------------------
 * 40 Hz -> 100 Hz reference interpolation of rt_mpc_qp (ref_interp.c): PRMPCClass::solve_AAA_inv_mod1 and
 * XGetSolution_position_mod3, RT/src/FastMPC/PRMPCClass.cpp:1170-1261,1344-1361.  inv row-major 4x4.
 * --------------------------------------------------------------------- */
void orc_interp_aaa_inv_mod(double dt, double inv[16]);
void orc_interp_position_mod3(const double inv[16], int nh, int t_end_footstep, int walktime, double dt_sample,
                              const double in1[3], const double in2[3], const double ref[3], const double ref2[3],
                              double *out);

/* Swing-foot roll / pitch reference of rt_mpc_qp (foot_rot.c): PRMPCClass::XGetSolution_Foot_rotation,
 * RT/src/FastMPC/PRMPCClass.cpp:2255-2380.  State = the members that persist between calls. */
typedef struct { int bjxx, bjx1; double Rr[15], Lr[15]; } orc_foot_rot_state;   /* 3x5 row-major angle members */
void orc_foot_rot_state_init(orc_foot_rot_state *s);
void orc_foot_rotation(const double tx[27], const double ts[27], const double td[27], const double footx[27], double footx_max,
                       double dt_mpc, int t_end_footstep, int nh, orc_foot_rot_state *s, int walktimex, double dt_sample, double *out);

void orc_foot_rotation_w(const double tx[27], const double ts[27], const double td[27], const double footx[27], double footx_max,
                         double dt_mpc, int t_end_footstep, int nh, int ncol, int *bjxx_io, int *bjx1_io, double *Rr, double *Lr,
                         int walktimex, double dt_sample, double *out);

/* Swing-foot generator of the 100 Hz node (rt_foot.c): PRMPCClass::Foot_trajectory_solve_mod2, RT/src/FastMPC/PRMPCClass.cpp:1756-2195.
 * State (doubles): [0,27) _ts | [27,54) [54,81) [81,108) _footxyz_real rows x y z | [108,135) _lift_height_ref | [135] _ry_left_right |
 * [136] _bjxx | [137] _bjx1 | then six arrays of nh + 2: _Rfootx _Rfooty _Rfootz _Lfootx _Lfooty _Lfootz. */
typedef struct { double dt, dt_mpc, tstep, tdsp_ratio, stepwidth0, lift_height; } orc_rt_foot_cfg;
void orc_rt_foot_cfg_default(orc_rt_foot_cfg *c);
int orc_rt_foot_state_doubles(int nh);
void orc_rt_foot_state_default(const orc_rt_foot_cfg *c, int nh, double *s);
void orc_rt_foot_traj(const orc_rt_foot_cfg *c, int nh, double *s, int j_indexx, int stopwalking, const double nrt[9], double *out);

/* Glue of the 100 Hz node (rt_glue.c): RT/src/gait_fast.cpp:113-372 (sample bookkeeping) and :505-746 (tick).  The four
 * PRMPCClass methods are reached through hooks, so the one restatement of the glue drives the unmodified class (golden
 * vectors) and the oracle restatements (the checker).  Windows are signal-major [2][nh] as orc_body_theta_mpc takes them. */
typedef struct {
    void (*mod3)(void *ctx, int nh, int walktime, double dt_sample, const double *in1, const double *in2, const double *ref, const double *ref2, double *out);
    void (*foot)(void *ctx, int nh, int j_index, int stop, const double *nrt9, double *out /* 6 (nh + 1) */);
    void (*rot)(void *ctx, int nh, int walktimex, double dt_sample, double *out /* 6 nh */);
    void (*body)(void *ctx, int nh, int i, const double *bodyangle_state, const double *zmp, const double *ang, const double *rfoot,
                 const double *lfoot, const double *comacc_z, double *out14);
    double (*tx_total)(void *ctx);          /* (int) _tx_total */
} orc_rt_hooks;
int orc_rt_node_doubles(int nh);
void orc_rt_node_default(int nh, double *node);
void orc_rt_node_tick(int nh, double *node, const orc_rt_hooks *hk, void *ctx, const double msg[100], int ctrl_flag,
                      const double bodyangle_state[4], double out100[100]);
void orc_rt_hooks_oracle(orc_rt_hooks *hk);
int orc_rt_ctx_bytes(void);
int orc_body_mpc_bytes(void);
void orc_rt_ctx_init(void *ctx_mem, int nh, double *foot_state, double *rot_state, orc_body_mpc *body);

/* ------------------------------------------------------------------------
 * Ground-reaction-force distribution of go1_servo's 1 kHz loop (Dynamiccclass,
 * GO1/src/whole_body_dynamics/dynmics_compute.cpp:55-427): closed-form split, the
 * 12-variable QP (12 equality columns of which the stance legs' are all-zero, 24
 * inequalities), joint torques.  Leg order FR, FL, RR, RL.
 * --------------------------------------------------------------------- */
typedef struct { double qp_alpha, qp_beta, qp_gama, fz_max, mu; } orc_grf_cfg;
void orc_grf_cfg_default(orc_grf_cfg *c);
void orc_grf_force_distribution(const double com_des[3], const double leg_des[12], const double F[6], int mode,
                                double yc, const double rfoot_des[3], const double lfoot_des[3],
                                double F_leg_ref[12]);
int orc_grf_force_opt(const orc_grf_cfg *c, const double base_p[3], const double leg_p[12], const double FT[6],
                      const double F_leg_guess[12], int mode, int right_support, double grf[12],
                      int *active, int *nactive, int *iters, int *qp_solution);
void orc_grf_joint_torques(const double Jaco[9], int swing, const double p_des[3], const double p_est[3],
                           const double pv_des[3], const double pv_est[3], const double F_ref[3],
                           const double grav[3], double swing_kp, double swing_kd, double tau[3]);

/* ------------------------------------------------------------------------
 * Signal filters of go1_servo (filters.c): butterworthLPF (GO1/src/Filter/butterworthLPF.cpp:82-121) and
 * ButterworthFilter::ForceFilter (butterworth_filter.cpp:37-69).  States are plain double arrays (layouts in filters.c).
 * --------------------------------------------------------------------- */
typedef struct { double b0, b1, b2, a1, a2, a; } orc_lpf_coef;
void orc_lpf_init(double fsampling, double fcutoff, orc_lpf_coef *c);
double orc_lpf_filter(const orc_lpf_coef *c, double state[5], double y);
double orc_force_filter(double state[6], double input);

/* ------------------------------------------------------------------------
 * Go1 leg kinematics (Kinematicclass, GO1/src/kinematics/Kinematics.cpp:29-304).
 * leg: 0 FR, 1 FL, 2 RR, 3 RL.  J is row-major 3x3 (the reference's Jacobian_kin
 * side channel after the call).  The IK functions return the number of Newton
 * updates applied.
 * --------------------------------------------------------------------- */
void orc_leg_fk(const double q[3], int leg, double pos[3], double J[9]);
void orc_leg_fk_g(const double bp[3], const double br[3], const double q[3], int leg, double pos[3], double J[9]);
int orc_leg_ik(const double pdes[3], const double qini[3], int leg, double q[3], double J[9]);
int orc_leg_ik_g(const double bp[3], const double br[3], const double pdes[3], const double qini[3], int leg,
                 double q[3], double J[9]);

#ifdef __cplusplus
}
#endif
#endif
