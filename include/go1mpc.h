/*
 * go1mpc.h -- C ABI of libgo1mpc.so: batched, B200-native (sm_100a, FP64)
 * implementation of the Go1 gait-planning MPC hot path of
 * jtdingx/quadrupedal_loco.  Plain pointers and sizes only; no C++ or torch
 * types cross this boundary.  Every entry point returns 0 on success or a
 * negative GO1MPC_E_* code; go1mpc_last_error() gives the text.  There is no
 * CPU fallback: without a CUDA device every compute call fails.
 *
 * Reference abbreviations (paths under the reference repository):
 *   RT  = unitree_ros/rt_mpc_qp/src        NLP = unitree_ros/mosek_nlp_kmp/src
 *   GO1 = unitree_ros/go1_rt_control/src
 *
 * The reference has no FFI: its seams are C++ classes.  Each function below
 * names the class method(s) it replaces; the C++ mirror of those classes that
 * binds to this ABI is in quadrupedal_loco_b200/host/ and INTEGRATION.md shows
 * the binding a maintainer of the reference would add.
 *
 * QP convention (RT/utils/EiQuadProg/EiQuadProg.hpp:15-32):
 *     min 0.5 x'Gx + g0'x   s.t.  CE'x + ce0 = 0,  CI'x + ci0 >= 0
 * Matrices are column-major double, one constraint per column of CE / CI.
 *
 * Batch layout ("field arrays of per-problem blocks"): every argument is one
 * array per field; problem b owns the contiguous block [b*stride, (b+1)*stride).
 * A warp (or thread) works on one problem, so a block is read with fully
 * coalesced 128-byte lines or one TMA bulk copy; blocks that are TMA sources
 * are documented as 16-byte aligned with an even number of doubles.
 * Pointers suffixed _d are DEVICE pointers owned by the caller; the *_host
 * variants take HOST pointers and do the staging copies inside the call.
 * `stream` is a cudaStream_t passed as void* (NULL = the handle's own stream).
 * A handle is thread-compatible, not thread-safe; use one per host thread/GPU.
 */
#ifndef GO1MPC_H
#define GO1MPC_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GO1MPC_VERSION_STRING "0.1.0"
#define GO1MPC_FOOTSTEPS 27          /* _footstepsnumber, RT/FastMPC/PRMPCClass.h:30 */
#define GO1MPC_BODY_NH_MAX 40

/* error codes (function return values) */
enum {
    GO1MPC_OK = 0,
    GO1MPC_E_INVALID = -1,     /* bad argument (size, NULL, alignment)        */
    GO1MPC_E_CUDA = -2,        /* CUDA runtime error, see go1mpc_last_error   */
    GO1MPC_E_NO_DEVICE = -3,   /* no usable CUDA device: there is no fallback */
    GO1MPC_E_UNSUPPORTED = -4  /* size outside what the kernels are built for */
};

/* per-problem solver status (status_d[b]); numerically = oracle's ORC_* */
enum {
    GO1MPC_QP_OK = 0,            /* converged                                  */
    GO1MPC_QP_NOT_PD = 1,        /* LLT failed: x untouched, cost = +inf       */
    GO1MPC_QP_INFEASIBLE = 2,    /* no step possible: cost = +inf              */
    GO1MPC_QP_ITER_CAP = 3,      /* safety cap hit (reference has no cap)      */
    GO1MPC_QP_NAN = 4,           /* NaN in x  (reference: solveQP() == false)  */
    GO1MPC_QP_EQ_DEPENDENT = 5,  /* dependent equalities: early return         */
    GO1MPC_QP_SKIPPED = -1       /* front-end gated this tick: no solve ran    */
};
/* iters_d[b*GO1MPC_ITERS + k]: outer passes (step 1), constraints added, dropped,
 * degenerate restarts, passes through step 2a, and the ALGORITHMIC flop count of the
 * dense reference algorithm along the path this problem took (SURVEY.md section 8d:
 * what the roofline figure uses; saturates at INT_MAX). */
#define GO1MPC_ITERS 6
enum { GO1MPC_IT_OUTER = 0, GO1MPC_IT_ADD = 1, GO1MPC_IT_DROP = 2, GO1MPC_IT_DEGEN = 3,
       GO1MPC_IT_L2A = 4, GO1MPC_IT_FLOPS = 5 };

/* Constants of the body-inclination MPC.  Defaults = the reference's
 * compile-time values: RT/Robotpara/robot_const_para_config.cpp:8-47,
 * RT/FastMPC/PRMPCClass.cpp:157-170,227-261,280-287. */
typedef struct {
    double dt_mpc;               /* 0.01   gait::dt_mpc_fast                   */
    double dt_slow;              /* 0.025  gait::dt_mpc_slow                   */
    double tstep;                /* 0.7    gait::t_period                      */
    double height_offset_time;   /* 1.0                                        */
    double g, mass, j_ini;       /* 9.8, 12, 12*0.1*0.1                        */
    double foot_length, foot_width; /* 0.02, 0.02                              */
    double theta_lim;            /* 10 deg in rad                              */
    double torque_lim;           /* 20 (the reference divides it by j_ini)     */
    double Rtheta, alphatheta, beltatheta, gama_zmp; /* 100, 10, 5e9, 5000     */
    double lamda[4];             /* state feedback gains (all 0 as shipped)    */
} Go1BodyMpcConfig;

/* Constants of the step-location / step-timing SQP.  Defaults = the reference's values:
 * NLP/NLP/NLPClass_sqp.cpp:212-214,244-258,273-286 (go1 weights), NLP/NLP/NLPClass.h:30-38,
 * NLP/NLPRTControl/NLPRTControlClass.cpp:35-42. */
typedef struct {
    double dt;                   /* 0.025  _dt                                  */
    double Wn;                   /* sqrt(g / hcom), hcom = 0.309458             */
    double ggg;                  /* 9.8                                         */
    double t_min, t_max;         /* 0.5, 1.0  step-period bounds                */
    double footx_max, footx_min; /* 0.15, -0.05                                 */
    double footx_vmax, footx_vmin, footy_vmax, footy_vmin;  /* 3, -2.875, 2, -1 */
    double comax_max, comax_min, comay_max, comay_min;      /* 5, -5, 6, -6     */
    double aax, aay, aaxv, aayv, bbx, bby, rr1, rr2;        /* 5e4 5e4 1e3 5e2 2e6 1e7 1e6 1e6 */
    double half_hip_width, foot_width;                       /* 0.12675, 0.03    */
    double lamda[4];             /* com x, vx, y, vy feedback gains (0 as shipped) */
    double hcom;                 /* 0.309458  RobotPara_Z_C - _height_offset    */
    int ext_height;              /* 0: CoM_height_solve on the device; 1: comz/comaz/comvz from in_d */
    int reserved;
    double stepwidth0;           /* 0.12675  _stepwidth(0) = stepwidthinput / 2 (NLPClass_sqp.cpp:66-67) */
    double lift_height;          /* 0.03     _lift_height (NLPRTControlClass.cpp:48), 0 for the last two steps */
} Go1StepMpcConfig;

typedef struct {
    Go1BodyMpcConfig body;
    Go1StepMpcConfig step;
    int qp_iter_cap_scale;       /* cap on step-2a passes = scale*(n+m+p)+50; default 20 */
    int reserved[7];
} Go1MpcConfig;

typedef struct go1mpc go1mpc_t;

const char *go1mpc_version(void);
int go1mpc_config_default(Go1MpcConfig *cfg);
/* device < 0 selects the current CUDA device */
int go1mpc_create(const Go1MpcConfig *cfg, int device, go1mpc_t **out);
void go1mpc_destroy(go1mpc_t *h);
const char *go1mpc_last_error(const go1mpc_t *h);
int go1mpc_device(const go1mpc_t *h);
/* number of kernels this handle has launched so far (bench's gpu_launches) */
long long go1mpc_launch_count(const go1mpc_t *h);
/* the handle's own cudaStream_t (as void*), e.g. to record timing events on it */
void *go1mpc_stream(const go1mpc_t *h);
/* multiprocessor count of the handle's device */
int go1mpc_sm_count(const go1mpc_t *h);
/* device-to-device copy on `stream` (NULL = the handle's): all per-instance state of this library
 * is plain SoA / record memory, so checkpointing or restoring a batch is a memcpy */
int go1mpc_copy_device_async(go1mpc_t *h, void *dst_d, const void *src_d, size_t bytes, void *stream);
/* The *_host_async entries order calls that share a device-resident buffer (planner state, body records) through a per-buffer
 * event; call this before freeing such a buffer (waits for its last user, drops the bookkeeping). */
int go1mpc_forget_buffer(go1mpc_t *h, const void *buf_d);
/* wait for the handle's stream and for every pipelined *_host_async call */
int go1mpc_synchronize(go1mpc_t *h);

/* ---------------------------------------------------------------------------
 * Generic dense QP batch.  Replaces, for B independent problems of one shape,
 *   QPBaseClass::solveQP            RT/QP/QPBaseClass.cpp:126-153
 *   Eigen::QP::solve_quadprog       RT/utils/EiQuadProg/EiQuadProg.cpp:493-513
 * Strides (doubles): G n*n, g0 n, CE n*p, ce0 p, CI n*m, ci0 m, x n, cost 1;
 * (ints): active m+p, nactive 1, iters GO1MPC_ITERS, status 1.  CE_d/ce0_d may be NULL
 * when p == 0.  x_d is in/out (left untouched when G is not PD, as the
 * reference does).  active_d/nactive_d/iters_d/cost_d may be NULL.
 * Limits: n <= 96, m + p <= 1024.
 * ------------------------------------------------------------------------ */
int go1mpc_qp_solve_batch(go1mpc_t *h, int n, int p, int m, int B,
                          const double *G_d, const double *g0_d,
                          const double *CE_d, const double *ce0_d,
                          const double *CI_d, const double *ci0_d,
                          double *x_d, double *cost_d,
                          int *active_d, int *nactive_d, int *iters_d, int *status_d,
                          void *stream);
int go1mpc_qp_solve_batch_host(go1mpc_t *h, int n, int p, int m, int B,
                               const double *G, const double *g0,
                               const double *CE, const double *ce0,
                               const double *CI, const double *ci0,
                               double *x, double *cost,
                               int *active, int *nactive, int *iters, int *status);

/* ---------------------------------------------------------------------------
 * Body-inclination MPC tick for B independent instances, horizon nh.
 * Replaces PRMPCClass::body_theta_mpc (condensation + QP + clamp + roll-out)
 *   RT/FastMPC/PRMPCClass.cpp:379-714, solve_body_rotation/Solve :799-849,
 *   Indexfind :716-738, with the model of Initialize :198-261,280-287.
 * One launch: condensation, the 2nh-variable / 12nh-row QP and the
 * post-processing are fused; the QP matrices never exist in HBM.
 *
 * in_d   [B][go1mpc_body_in_stride(nh)] doubles, 16-byte aligned, per problem:
 *          [0,27)      tx        step-cycle start times (_tx)
 *          [27]        tick      i, stored as a double (exact integer)
 *          [28,32)     theta     (thetaxk0, thetaxk1, thetayk0, thetayk1)
 *          [32,36)     bodyangle_state (only used with non-zero lamda)
 *          [36,36+2nh) x_warm    _V_ini (kept when G is not PD / tick gated)
 *          then 9 rows of nh:  zmp_x, zmp_y, ang_x, ang_y, rfoot_x, rfoot_y,
 *                              lfoot_x, lfoot_y, comacc_z
 *          (+1 pad double when the count is odd)
 * out_d  [B][go1mpc_body_out_stride(nh)] doubles, in/out (a gated tick keeps
 *        its out14, as the reference returns its stale members):
 *          [0,14)      out14     = the Vec14 body_theta_mpc returns (:696-709)
 *          [14,18)     theta     updated state
 *          [18,18+2nh) x         _V_ini after the tick
 *          [18+2nh]    cost
 *          (+pad to an even count)
 * diag_d [B][go1mpc_body_diag_stride(nh)] ints (may be NULL):
 *          [0] status  [1] nactive  [2,6) iters  [6] bjx1  [7] bjx2
 *          [8] passes through step 2a  [9] algorithmic flops (GO1MPC_IT_FLOPS)
 *          [10, 10+2nh) final active set (constraint indices, Appendix E of SURVEY.md)
 * Supported nh: 3..40 (the reference compiles nh = 4).
 * ------------------------------------------------------------------------ */
int go1mpc_body_in_stride(int nh);
int go1mpc_body_out_stride(int nh);
int go1mpc_body_diag_stride(int nh);
int go1mpc_body_mpc_step_batch(go1mpc_t *h, int nh, int B,
                               const double *in_d, double *out_d, int *diag_d,
                               void *stream);
/* Diagnostic: number of instances the roll/pitch-split kernel handed to the combined-solve kernel
 * (infeasible / degenerate / non-converged halves) since the handle was created.  Synchronises the device. */
int go1mpc_body_handover_total(go1mpc_t *h, long long *total);
/* Diagnostic: warps of the body solve kernel that left through its defensive iteration guard (0 unless
 * there is a bug).  Synchronises the device. */
int go1mpc_body_guard_trips(go1mpc_t *h, long long *total);
int go1mpc_body_mpc_step_batch_host(go1mpc_t *h, int nh, int B,
                                    const double *in, double *out, int *diag);
/* Pipelined host entries: enqueue H2D copy, kernel and D2H copy on one of the handle's eight
 * internal lanes (consecutive calls use consecutive lanes, so the copies of one batch overlap
 * the kernels of its neighbours) and return at once; go1mpc_synchronize() waits for all of
 * them.  Host buffers should be pinned and must stay untouched until then.  The planner state
 * of the step-timing tick stays RESIDENT ON THE DEVICE (state_d / state_out_d are device
 * pointers, as in go1mpc_step_timing_step_batch); only the per-tick inputs and results move.
 * Calls that write a state buffer an earlier pipelined call wrote are ordered after it. */
int go1mpc_body_mpc_step_batch_host_async(go1mpc_t *h, int nh, int B,
                                          const double *in, double *out, int *diag);
int go1mpc_step_timing_step_batch_host_async(go1mpc_t *h, int n_sqp, int B, const int *tick,
                                             const double *state_d, double *state_out_d,
                                             const double *in, double *out, int *diag);
/* Pipelined body tick with the slow part of the record RESIDENT ON THE DEVICE, as the planner state is:
 * what PRMPCClass::body_theta_mpc takes per call (RT/FastMPC/PRMPCClass.cpp:379-395) moves, what the class
 * keeps as members (_tx :174-178, _V_ini :258-261, the stale results a gated tick returns) stays in HBM.
 *   tx_d     device [B][28] doubles: tx[27] + 1 pad (the caller rewrites a row when the planner re-times a step)
 *   out_d    device [B][go1mpc_body_out_stride(nh)], in/out: the output records of go1mpc_body_mpc_step_batch;
 *            the tick reads its warm start x (and, when gated, its out14) from here and writes its results back
 *   tick_in  host [B][go1mpc_body_tick_in_stride(nh)]: tick | theta(4) | bodyangle_state(4) | 9 rows of nh
 *            (= record doubles [27,36) and [36+2nh, 36+11nh) of the full format; + pad to an even count)
 *   tick_out host [B][GO1MPC_BODY_TICK_OUT]: out14 | theta(4) | cost | 0
 *   diag     host [B][go1mpc_body_diag_stride(nh)] or NULL
 * Results are bit-identical to the full-record entries.  Calls on the same out_d are ordered. */
#define GO1MPC_BODY_TICK_OUT 20
int go1mpc_body_tick_in_stride(int nh);
int go1mpc_body_mpc_step_batch_resident_host_async(go1mpc_t *h, int nh, int B,
                                                   const double *tx_d, double *out_d,
                                                   const double *tick_in, double *tick_out, int *diag);

/* Model matrices the handle condenses with, for inspection/tests (host
 * buffers, column-major): pps,pvs nh x 2; ppu,pvu,ppu_2,pvu_2 nh x nh.
 * Replaces PRMPCClass::Matrix_ps / Matrix_pu, RT/FastMPC/PRMPCClass.cpp:741-796. */
int go1mpc_body_model(go1mpc_t *h, int nh, double *pps, double *pvs,
                      double *ppu, double *pvu, double *ppu_2, double *pvu_2);
/* default step table _tx of PRMPCClass::Initialize (:174-178), 27 doubles */
int go1mpc_body_default_tx(go1mpc_t *h, double *tx27);

/* ---------------------------------------------------------------------------
 * Step-location / step-timing SQP tick (40 Hz planner) for B independent instances.
 * Replaces NLPClass::step_timing_opti_loop  NLP/NLP/NLPClass_sqp.cpp:693-1102 with
 *   step_timing_object_function :1144-1173, step_timing_constraints :1175-1458,
 *   solve_stepping_timing/Solve :1613-1653 (QP n=4, p=1, m=24), Indexfind :1105-1141.
 * One launch runs n_sqp SQP iterations (reference: 3; 1..5 supported), the write-back of
 * step length / width / period and of the step tables, the LIPM roll-out of samples
 * i..i+2, the feedback blend, the integer step indices and CoM_height_solve (:2361-2473,
 * the 6th-order vertical CoM polynomial).
 *
 * Layout: STRUCTURE OF ARRAYS, element-major / batch-minor: field f of instance b is
 * at [f*B + b] (one thread per instance: every access of a warp is coalesced).
 * tick_d  [B] ints         i of the reference (>= 1); an instance with tick < 1 is skipped (nothing read or written)
 * state_d [202][B] doubles, the planner state before the tick; state_out_d receives the
 *         state after it and may be the same buffer (in-place: only changed fields are
 *         written).  Fields:
 *           [0,27) ts   [27,54) tx   [54,81) footx_ref  [81,108) footy_ref
 *           [108,135) footz_ref  [135,162) Lxx_ref  [162,189) Lyy_ref
 *           [189,195) com x,vx,ax,y,vy,ay _feed at tick i-1
 *           [195,199) _Vari_ini.col(i-1) = (Lx, Ly, cosh(w T), sinh(w T))
 *           [199,201) _comvx_endref, _comvy_endref   [201] _bjx1 of the previous tick (exact integer)
 * in_d    [20][B] doubles: [0,6) estimated com x,vx,ax,y,vy,ay  [6,8) right foot x,y
 *           [8,10) left foot x,y  [16,19) Zsc(i..i+2) (terrain height under the CoM)
 *           [10,13) comz(i..i+2)  [13,16) comaz  [19] comvz(i): read only when cfg.step.ext_height != 0
 * out_d   [38][B] doubles = the Vec38 the reference returns (:1048-1090)
 * diag_d  [60][B] ints (may be NULL): [0] period index (_periond_i; -1 = time beyond the
 *           step table, nothing written)  [1] k_yu  [2] bjxx  [3] bjx1  [4] QPs solved;
 *           then per SQP iteration q < 5, at 5 + 11 q: status (-1 = no solve), nactive,
 *           iters[4], active set[5] (slot 0 = -1 is the equality)
 * ------------------------------------------------------------------------ */
#define GO1MPC_STEP_STATE_DOUBLES 202
#define GO1MPC_STEP_IN_DOUBLES 20
#define GO1MPC_STEP_OUT_DOUBLES 38
#define GO1MPC_STEP_DIAG_INTS 60
int go1mpc_step_timing_step_batch(go1mpc_t *h, int n_sqp, int B, const int *tick_d,
                                  const double *state_d, double *state_out_d,
                                  const double *in_d, double *out_d,
                                  int *diag_d, void *stream);
int go1mpc_step_timing_step_batch_host(go1mpc_t *h, int n_sqp, int B, const int *tick,
                                       double *state, const double *in, double *out, int *diag);
/* Initial step tables of one planner (host buffer, 202 doubles, instance-major):
 * NLPClass::FootStepInputs :51-75 and Initialize :131-206 for the given step length,
 * width and height (reference: 0.075, 0.2535, 0) and period tstep (0.7). */
int go1mpc_step_default_state(go1mpc_t *h, double steplength, double stepwidth,
                              double stepheight, double tstep, double *state201);

/* ---------------------------------------------------------------------------
 * Go1 leg kinematics for B independent legs.  Replaces Kinematicclass
 * (GO1/kinematics/Kinematics.h:49-60, Kinematics.cpp:29-304):
 *   go1mpc_leg_fk_batch  -> Forward_kinematics (body_p_d == NULL) / Forward_kinematics_g
 *   go1mpc_leg_ik_batch  -> Inverse_kinematics (body_p_d == NULL: <= 10 updates, stop
 *                           max(dq) < 1e-4 without abs, as the reference) / Inverse_kinematics_g
 *                           (<= 15 updates, stop |dp|^2 <= 1e-6); lamda = 0.5
 * The reference returns the Jacobian through the member Jacobian_kin read after each
 * call (servo.cpp:734-741,1038-1051); here it is the output array jac_d.
 * Layout: structure of arrays, [f*B + b].  q_d / qini_d / pdes_d / pos_d / body_p_d /
 * body_r_d (roll, pitch, yaw): [3][B] doubles; jac_d: [9][B] doubles, row-major 3x3, may be
 * NULL; leg_d: [B] ints, 0 FR, 1 FL, 2 RR, 3 RL; iters_d: [B] ints (Newton updates), may be NULL.
 * ------------------------------------------------------------------------ */
int go1mpc_leg_fk_batch(go1mpc_t *h, int B, const double *q_d, const int *leg_d,
                        const double *body_p_d, const double *body_r_d,
                        double *pos_d, double *jac_d, void *stream);
int go1mpc_leg_ik_batch(go1mpc_t *h, int B, const double *pdes_d, const double *qini_d, const int *leg_d,
                        const double *body_p_d, const double *body_r_d,
                        double *q_d, double *jac_d, int *iters_d, void *stream);
int go1mpc_leg_fk_batch_host(go1mpc_t *h, int B, const double *q, const int *leg,
                             const double *body_p, const double *body_r, double *pos, double *jac);
int go1mpc_leg_ik_batch_host(go1mpc_t *h, int B, const double *pdes, const double *qini, const int *leg,
                             const double *body_p, const double *body_r, double *q, double *jac, int *iters);

/* ---------------------------------------------------------------------------
 * Swing-foot trajectory of the step planner for B instances; run after the step-timing tick
 * of the same index.  Replaces NLPClass::Foot_trajectory_solve_mod2 (NLP/NLP/NLPClass_sqp.cpp:
 * 2039-2358) and solve_AAA_inv2 (:3633-3645); the stop-walking branch is go1mpc_foot_trajectory_stop_batch.
 * SoA layout [f*B + b].  Instances with tick < 1 are skipped.
 * state_d [202][B]  planner state AFTER go1mpc_step_timing_step_batch;  out38_d [38][B] its output
 * foot_d  [32][B] in/out, the window of the reference's whole-walk foot arrays:
 *           [0,6) R xyz, L xyz at tick j-1   [6,12) what the arrays hold at j   [12,18) at j-2
 *           [18,24) at j-3   [24,30) frozen at the step's start   [30] step start index the freeze
 *           belongs to (-1 none)   [31] _ry_left_right.  Initial value: go1mpc_foot_default_state.
 * out18_d [18][B] = the Vec18 of the reference (R xyz, L xyz | velocities | accelerations)
 * right_support_d [B] ints (0 left support, 1 right support, 2 double support), may be NULL
 * ------------------------------------------------------------------------ */
#define GO1MPC_FOOT_STATE_DOUBLES 32
#define GO1MPC_FOOT_OUT_DOUBLES 18
int go1mpc_foot_trajectory_batch(go1mpc_t *h, int B, const int *tick_d, const double *state_d,
                                 const double *out38_d, double *foot_d, double *out18_d,
                                 int *right_support_d, void *stream);
int go1mpc_foot_trajectory_batch_host(go1mpc_t *h, int B, const int *tick, const double *state,
                                      const double *out38, double *foot, double *out18, int *right_support);
int go1mpc_foot_default_state(go1mpc_t *h, double *foot32);
/* The same with the stop-walking branch (:2043-2048): lift0_d [B] doubles, in/out -- the first step index whose lift height is
 * zeroed (initially GO1MPC_FOOTSTEPS = none); stop_d [B] doubles or NULL -- the caller's _stopwalking flag (non-zero = set).
 * A tick beyond _t_end_footstep (go1mpc_nlp_t_end_footstep) acts like a stop, as in the reference. */
int go1mpc_foot_trajectory_stop_batch(go1mpc_t *h, int B, const int *tick_d, const double *state_d, const double *out38_d,
                                      double *foot_d, double *out18_d, int *right_support_d, double *lift0_d,
                                      const double *stop_d, void *stream);
int go1mpc_foot_trajectory_stop_batch_host(go1mpc_t *h, int B, const int *tick, const double *state, const double *out38,
                                           double *foot, double *out18, int *right_support, double *lift0, const double *stop);

/* ---------------------------------------------------------------------------
 * Servo kinematics tick for B robots (cfg5's fused leg IK / Jacobian stage).  Replaces the
 * leg mapping and the four Inverse_kinematics_g calls of go1_servo's 1 kHz loop,
 * GO1/servo_control/servo.cpp:935-1051: foot target of each leg = its homing position + the
 * planner's virtual right / left foot displacement -+ half hip width, the virtual foot chosen
 * by gait_mode (101 pace, 102 trot, 103 gallop; servo.h:100-102), body pose = (com x, com y *
 * y_offset, com z; roll, pitch, yaw), initial guess = the previous joint angles.
 * SoA [f*B + b]: com_d, theta_d, rfoot_d, lfoot_d [3][B]; homing_d [12][B] and q_d [12][B]
 * (in/out) leg-major in the order FR, FL, RR, RL; jac_d [36][B] (row-major 3x3 per leg),
 * foot_des_d [12][B], iters_d [4][B] may be NULL.
 * ------------------------------------------------------------------------ */
int go1mpc_servo_kin_tick_batch(go1mpc_t *h, int B, int gait_mode, double y_offset,
                                const double *com_d, const double *theta_d,
                                const double *rfoot_d, const double *lfoot_d, const double *homing_d,
                                double *q_d, double *jac_d, double *foot_des_d, int *iters_d, void *stream);

/* ---------------------------------------------------------------------------
 * Fused control tick for B robots: ONE call enqueues, on one stream and with every
 * intermediate resident on the device,
 *   1. the step-location / step-timing SQP tick   (= go1mpc_step_timing_step_batch, state updated in place)
 *   2. the planner's swing-foot trajectory        (= go1mpc_foot_trajectory_batch on the new state / out38)
 *   3. the body-inclination MPC tick               (= go1mpc_body_mpc_step_batch)
 *   4. the servo kinematics tick                   (= go1mpc_servo_kin_tick_batch) with the body pose taken from the
 *      ticks above: position = planner CoM (out38 rows 0..2), roll / pitch = body-MPC angles (out record [0], [1]),
 *      yaw 0; virtual right / left foot = swing-foot positions (out18 rows 0..2 / 3..5)
 *   5. optionally the GRF QP and the torque map    (= go1mpc_grf_force_opt_batch, go1mpc_grf_joint_torques_batch on stage 4's
 *      Jacobians and the QP's forces: Dynamiccclass::force_opt / compute_joint_torques, servo.cpp:1200-1243)
 * i.e. NLPClass::step_timing_opti_loop + Foot_trajectory_solve_mod2 -> PRMPCClass::body_theta_mpc -> the leg mapping
 * and four Inverse_kinematics_g of GO1/servo_control/servo.cpp:935-1051 without a host round trip.  Results are
 * bit-identical to calling the four entry points in that order.  All pointers are device pointers with the layouts
 * of the individual entry points; servo_theta_d [3][B] is scratch the call fills (roll, pitch, yaw).
 * ------------------------------------------------------------------------ */
typedef struct {
  int n_sqp;                    /* 1: planner */
  const int *tick_d;
  double *step_state_d;         /*    [202][B], in place */
  const double *step_in_d;      /*    [20][B] */
  double *out38_d;              /*    [38][B] */
  int *step_diag_d;             /*    [60][B] or NULL */
  double *foot_d;               /* 2: [32][B] in/out */
  double *out18_d;              /*    [18][B] */
  int *right_support_d;         /*    [B] or NULL */
  int nh;                       /* 3: body MPC */
  const double *body_in_d;      /*    [B][go1mpc_body_in_stride(nh)] */
  double *body_out_d;           /*    [B][go1mpc_body_out_stride(nh)] in/out */
  int *body_diag_d;             /*    [B][go1mpc_body_diag_stride(nh)] or NULL */
  int gait_mode;                /* 4: servo kinematics (101 pace, 102 trot, 103 gallop) */
  double y_offset;
  const double *homing_d;       /*    [12][B] */
  double *q_d;                  /*    [12][B] in/out */
  double *jac_d;                /*    [36][B] or NULL */
  double *foot_des_d;           /*    [12][B] or NULL */
  int *ik_iters_d;              /*    [4][B] or NULL */
  double *servo_theta_d;        /*    [3][B] scratch (filled by the call) */
  const double *grf_in_d;       /* 5: optional (NULL skips it): force QP records [B][48] of go1mpc_grf_force_opt_batch */
  double *grf_out_d;            /*    [B][16] */
  int *grf_diag_d;              /*    [B][32] or NULL */
  const int *swing_d;           /*    torque map (tau_d NULL skips it): [4][B] */
  const double *p_des_d, *p_est_d, *pv_des_d, *pv_est_d;   /* [12][B] each */
  double *tau_d;                /*    [12][B]: -J' f + gravity term from stage 4's jac_d and the QP's forces */
} Go1FusedTick;
int go1mpc_fused_tick_batch(go1mpc_t *h, int B, const Go1FusedTick *t, void *stream);

/* ---------------------------------------------------------------------------
 * Ground-reaction-force distribution of go1_servo's 1 kHz loop for B robots.  Replaces
 * Dynamiccclass (GO1/whole_body_dynamics/dynmics_compute.cpp):
 *   go1mpc_grf_force_distribution_batch -> force_distribution :141-261 (gait_mode 101 / 102)
 *   go1mpc_grf_force_opt_batch          -> force_opt :265-373 + solve_grf_opt :387-427: the
 *       12-variable QP (alpha 1e4, beta 1e3, gama 10, fz_max 160, mu 0.25 of :55-66) with 12
 *       equality columns (identity on the legs that must carry no force, all-zero -- and
 *       skipped by the solver -- on the others) and 24 inequalities; on a NaN solution the
 *       closed-form guess is returned, as the reference does.
 * force_opt records are instance-major:
 *   in_d  [B][48]: base_p 3 | leg_p 12 (FR, FL, RR, RL xyz) | FT_total_des 6 | F_leg_guess 12 |
 *                  previous grf_opt 12 | gait_mode | right_support | pad
 *   out_d [B][16]: grf_opt 12 | cost | qp_solution (1/0) | pad 2
 *   diag_d [B][32] ints (may be NULL): status, nactive, iters[4], qp_solution, 0, active set[24]
 * force_distribution is SoA [f*B + b]: com_des [3][B], leg_des [12][B], F_force_des [6][B]
 * (L xyz, R xyz), rfoot_des / lfoot_des [3][B] -> F_leg_ref [12][B] (= F_leg_guess).
 * ------------------------------------------------------------------------ */
int go1mpc_grf_force_opt_batch(go1mpc_t *h, int B, const double *in_d, double *out_d, int *diag_d, void *stream);
int go1mpc_grf_force_distribution_batch(go1mpc_t *h, int B, int gait_mode, double y_coefficient,
                                        const double *com_des_d, const double *leg_des_d, const double *F_force_des_d,
                                        const double *rfoot_des_d, const double *lfoot_des_d,
                                        double *F_leg_ref_d, void *stream);
/* Joint torques of the four legs.  Replaces Dynamiccclass::compute_joint_torques :109-138 as go1_servo calls it once per
 * leg (GO1/servo_control/servo.cpp:1232-1243, leg_number 0..3 = FR, FL, RR, RL):
 *   tau = -Jaco' * w + gravity_compensate(:, leg),   w = swing_kp (p_des - p_est) + swing_kd (pv_des - pv_est) for a
 *   swing leg (swing_kp 1, swing_kd 0.01, :37-38), the leg's column of F_leg_ref for a stance leg.
 * SoA [f*B + b]: jac_d [36][B] (row-major 3x3 per leg, the layout go1mpc_servo_kin_tick_batch / go1mpc_leg_ik_batch
 * write), swing_d [4][B] ints, p_des_d / p_est_d / pv_des_d / pv_est_d [12][B], tau_d [12][B].  F_leg_ref element k of
 * robot b is read at F_leg_ref_d[k * F_elem_stride + b * F_robot_stride]: (B, 1) for the SoA output of
 * go1mpc_grf_force_distribution_batch, (1, 16) for the out records of go1mpc_grf_force_opt_batch. */
int go1mpc_grf_joint_torques_batch(go1mpc_t *h, int B, const double *jac_d, const int *swing_d,
                                   const double *p_des_d, const double *p_est_d,
                                   const double *pv_des_d, const double *pv_est_d,
                                   const double *F_leg_ref_d, long long F_elem_stride, long long F_robot_stride,
                                   double *tau_d, void *stream);

/* Synchronous host-buffer forms of the three entries above (same layouts, host pointers; what the C++ mirror of
 * Dynamiccclass in host/go1mpc.hpp calls with B = 1). */
int go1mpc_grf_force_opt_batch_host(go1mpc_t *h, int B, const double *in, double *out, int *diag);
int go1mpc_grf_force_distribution_batch_host(go1mpc_t *h, int B, int gait_mode, double y_coefficient,
                                             const double *com_des, const double *leg_des, const double *F_force_des,
                                             const double *rfoot_des, const double *lfoot_des, double *F_leg_ref);
int go1mpc_grf_joint_torques_batch_host(go1mpc_t *h, int B, const double *jac, const int *swing,
                                        const double *p_des, const double *p_est,
                                        const double *pv_des, const double *pv_est,
                                        const double *F_leg_ref, long long F_elem_stride, long long F_robot_stride,
                                        double *tau);

/* ---------------------------------------------------------------------------
 * 40 Hz -> 100 Hz reference interpolation for B items (one item = one 3-vector quantity of one robot).  Replaces
 * PRMPCClass::XGetSolution_position_mod3 (RT/FastMPC/PRMPCClass.cpp:1170-1261, with _AAA_inv_mod of
 * solve_AAA_inv_mod1 :1344-1361) as rt_mpc_qp calls it per quantity (RT/gait_fast.cpp:131-138): the cubic through
 * four consecutive 40 Hz samples, evaluated at walktime * dt_sample + jx * dt_sample, jx = 0..nh-1.
 *   walktime_d [B] ints; samples_d [12][B]: in1 xyz | in2 xyz | ref xyz | ref2 xyz (SoA [f*B + b])
 *   out_d [9 + 3 (nh - 1)][B]: position, velocity, acceleration at jx = 0, then the positions at jx = 1..nh-1
 *   (the reference's Vec21 at its nh = 4); all zero once walktime > _t_end_footstep.
 * Parity (tests/test_zz_ref_interp.py, B200): 1e-9 relative to the pinned oracle (oracle/ref_interp.c; the reference
 * evaluates the monomials with libm pow, the kernel with a correctly rounded integer power), 1e-12 absolute over the
 * reference's own walktime range (count_inteplotation = 1..2, RT/gait_fast.cpp:113-131).
 * ------------------------------------------------------------------------ */
int go1mpc_ref_interp_batch(go1mpc_t *h, int B, int nh, const int *walktime_d, double dt_sample,
                            const double *samples_d, double *out_d, void *stream);
/* Host part, needs no device: _AAA_inv_mod (row-major 4x4) and _t_end_footstep (:179) for a configuration
 * (NULL = defaults); either output may be NULL. */
int go1mpc_ref_interp_model(const Go1MpcConfig *cfg, double *inv16, int *t_end_footstep);

/* ---------------------------------------------------------------------------
 * One control tick for B robots with HOST inputs and COMPACT results -- the call a controller farm makes per tick:
 * planner tick (= go1mpc_step_timing_step_batch) and body-inclination MPC tick on the device-resident records
 * (= go1mpc_body_mpc_step_batch_resident_host_async) on ONE caller stream, then a 12-double result row per robot.
 * What the reference classes keep as members stays in HBM (planner state, _tx, _V_ini / stale results); what their
 * tick methods take as arguments moves up per call (tick, the 20 -- or 10, see step_in_rows -- planner inputs, the 9+9nh
 * body-tick doubles); what
 * comes down is GO1MPC_COMPACT_DOUBLES per robot instead of Vec38 + Vec14 + diagnostics:
 *   [0,3) CoM x y z (Vec38 0..2)   [3,5) body roll, pitch   [5,7) body torques (Vec14 0..3)
 *   [7,9) next footstep x, y (Vec38 29, 31)   [9] step period (Vec38 35)
 *   [10] status of the planner's last SQP solve (-1: none ran)   [11] body QP status
 * compact_d is a caller-owned DEVICE buffer [B][12] (16-byte aligned; e.g. the send buffer of the once-per-batch gather
 * to rank 0); `compact` (host, may be NULL) receives a copy.  out38 / step_diag / body_diag (host, may be NULL) receive
 * the full results.  step_state_src_d (may be NULL): run the planner tick out of place from this pristine state.
 * Everything is enqueued on `stream`; staging is per stream and owned by the handle.
 * ------------------------------------------------------------------------ */
#define GO1MPC_COMPACT_DOUBLES 12
typedef struct {
  int n_sqp, nh;
  const int *tick;                /* host [B] */
  const double *step_in;          /* host [20][B] */
  const double *body_tick_in;     /* host [B][go1mpc_body_tick_in_stride(nh)] */
  const double *step_state_src_d; /* device [202][B] or NULL */
  double *step_state_d;           /* device [202][B], state after the tick */
  const double *tx_d;             /* device [B][28] */
  double *body_out_d;             /* device [B][go1mpc_body_out_stride(nh)], in/out */
  double *compact_d;              /* device [B][12] */
  double *compact;                /* host [B][12] or NULL */
  double *out38;                  /* host [38][B] or NULL */
  int *step_diag;                 /* host [60][B] or NULL */
  int *body_diag;                 /* host [B][go1mpc_body_diag_stride(nh)] or NULL */
  int step_in_rows;               /* leading rows of step_in that are uploaded: 0 or 20 = all; 10 = the sensor rows only
                                   * (estimated CoM state, foot locations; step_in is then host [10][B]).  Rows 10..19 -- the
                                   * external CoM-height samples (read only with cfg.step.ext_height) and the terrain heights
                                   * Zsc -- then read as zero: flat ground, the reference's own configuration */
} Go1ControlTick;
int go1mpc_control_tick_host_async(go1mpc_t *h, int B, const Go1ControlTick *t, void *stream);
/* The pack kernel alone (device pointers; step_diag_d / body_diag_d may be NULL). */
int go1mpc_pack_compact_batch(go1mpc_t *h, int B, int nh, const double *out38_d, const int *step_diag_d,
                              const double *body_out_d, const int *body_diag_d, double *compact_d, void *stream);

/* ---------------------------------------------------------------------------
 * Once-per-batch gather of result rows to rank 0 over NVLink / NVSwitch peer memory, without a collective.
 * Rank 0 owns `slots` blocks [world][rows_per_rank][row_doubles]; it exports an IPC blob the host program hands to the
 * other ranks (one process per GPU); a peer maps the memory and lets its kernels write its rows straight into rank 0's
 * block: go1mpc_gather_dest() is a valid compact_d for go1mpc_control_tick_host_async.  Per use of a slot, all on a stream:
 *   peer:   acquire -> [tick writing to dest] -> publish
 *   rank 0: [tick writing to dest] -> publish -> wait_all -> [read go1mpc_gather_block, e.g. D2H] -> release
 * acquire / wait_all are single-thread device-side waits on flags in rank 0's memory (about 10 s time-out -> go1mpc_gather_status);
 * no rank ever waits for another on the host.  Every rank must use a slot the same number of times, in the same order.
 * ------------------------------------------------------------------------ */
typedef struct go1mpc_gather go1mpc_gather_t;
#define GO1MPC_GATHER_BLOB_BYTES 128
int go1mpc_gather_create(int device, int rank, int world, int slots, int rows_per_rank, int row_doubles, go1mpc_gather_t **out);
void go1mpc_gather_destroy(go1mpc_gather_t *g);
int go1mpc_gather_export(go1mpc_gather_t *g, void *blob, int blob_bytes);          /* rank 0 */
int go1mpc_gather_import(go1mpc_gather_t *g, const void *blob, int blob_bytes);    /* peers */
double *go1mpc_gather_dest(go1mpc_gather_t *g, int slot);
const double *go1mpc_gather_block(go1mpc_gather_t *g, int slot);                   /* rank 0 */
int go1mpc_gather_acquire(go1mpc_gather_t *g, int slot, void *stream);
int go1mpc_gather_publish(go1mpc_gather_t *g, int slot, void *stream);
int go1mpc_gather_wait_all(go1mpc_gather_t *g, int slot, void *stream);            /* rank 0 */
int go1mpc_gather_release(go1mpc_gather_t *g, int slot, void *stream);             /* rank 0 */
int go1mpc_gather_status(go1mpc_gather_t *g, int *status);

/* ---------------------------------------------------------------------------
 * Stream ordering and CUDA-graph capture.  Every *_batch entry only enqueues work on the stream it is given and
 * allocates nothing once that stream has been used with the same batch size, so a sequence of ticks dealt over several
 * streams can be captured once and replayed with one launch (launch-bound loops at small batches):
 *   go1mpc_graph_capture_begin(h, s0); go1mpc_stream_wait(h, s1, s0); ...ticks on s0, s1...;
 *   go1mpc_stream_wait(h, s0, s1); go1mpc_graph_capture_end(h, s0, &g); go1mpc_graph_launch(h, g, s0);
 * go1mpc_stream_wait makes `waiter` wait for the work enqueued so far on `signaller` (NULL = the handle's stream).
 * ------------------------------------------------------------------------ */
int go1mpc_stream_wait(go1mpc_t *h, void *waiter, void *signaller);
int go1mpc_graph_capture_begin(go1mpc_t *h, void *stream);
int go1mpc_graph_capture_end(go1mpc_t *h, void *stream, void **graph_exec_out);
int go1mpc_graph_launch(go1mpc_t *h, void *graph_exec, void *stream);
int go1mpc_graph_destroy(go1mpc_t *h, void *graph_exec);

/* ---------------------------------------------------------------------------
 * Signal filters of go1_servo's 1 kHz loop for B robots x C channels.  Replaces butterworthLPF::filter
 * (GO1/Filter/butterworthLPF.cpp:104-121; init :82-100 = go1mpc_lpf_coefficients) -- 28 objects on slots of the /MPC/Gait
 * message, servo.cpp:579-610,898-931 -- and ButterworthFilter::ForceFilter (GO1/Filter/butterworth_filter.cpp:37-69).
 * coef      host, [C][6]: b0 b1 b2 a1 a2 a per channel (C <= 32 per call)
 * in_d      [rows][B] SoA; channel c filters row in_rows_d[c] (device ints) or row c when in_rows_d is NULL -- so the channels
 *           can be picked straight out of a [100][B] message buffer
 * state_d   [5][C][B] SoA, in/out: call counter, y_p, y_pp, x_p, x_pp of every filter object; zero-initialise
 *           (force filter: [6][C][B]: count, raw[2], filtered[3])
 * out_d     [C][B]
 * Bit-identical to the reference classes (tests/test_filters.py).
 * ------------------------------------------------------------------------ */
int go1mpc_lpf_coefficients(double fsampling, double fcutoff, double *coef6);
int go1mpc_lpf_batch(go1mpc_t *h, int B, int C, const double *coef, const double *in_d, const int *in_rows_d, double *state_d,
                     double *out_d, void *stream);
int go1mpc_force_filter_batch(go1mpc_t *h, int B, int C, const double *in_d, double *state_d, double *out_d, void *stream);

/* ---------------------------------------------------------------------------
 * The 40 Hz planner node of mosek_nlp_kmp for B robots: one call of NLPRTControlClass::WalkingReactStepping per robot,
 * /MPC/Gait message out.  Replaces NLP/NLPRTControl/NLPRTControlClass.cpp:191-396 (squat / walk / over / idle branches and the
 * 100-slot message layout :288-392), StartWalking / StopWalking :400-432, rt_nlp_gait :436-596 with everything it calls:
 *   NLPClass::X_CoM_position_squat (NLP/NLP/NLPClass_sqp.cpp:2958-3015), step_timing_opti_loop + CoM_height_solve,
 *   Foot_trajectory_solve_mod2 with its stop-walking branch, Zmp_distributor / zmp_interpolation / Force_torque_calculate
 *   (:3650-3897).
 * state_d     [go1mpc_nlp_node_state_doubles()][B] SoA, in/out: rows [0,202) the planner state, [202,234) the swing-foot
 *             window, then the ZMP ring and the node's members (go1mpc_nlp_node_default_state gives one robot's column)
 * walkdtime_d [B] ints: the caller's tick counter (the reference's walkdtime, 1, 2, 3 ...; the walk's ticks must be consecutive)
 * start_d     [B] ints or NULL (= 1): start_mpc;   cmd_d [B] ints or NULL: 1 = StopWalking(), 2 = StartWalking() before the tick
 * rfoot_fb_d / lfoot_fb_d  [3][B] or NULL (= 0): measured foot locations.  The reference ignores its estimated-state argument
 *             (rt_nlp_gait passes the all-zero member _estimated_state, :446), so there is none here.
 * msg_d       [100][B] SoA: the message (slot 98 = 0)
 * Six launches on `stream`.  Parity: tests/test_gpu_nlp_node.py (cfg1 and stop / restart scripts of the unmodified class).
 * ------------------------------------------------------------------------ */
/* rows of the node state a host may read or set between ticks (exact integers / flags stored as doubles) */
#define GO1MPC_NLP_ROW_STOP 365          /* _stop_walking */
#define GO1MPC_NLP_ROW_START_AGAIN 366   /* _start_walking_again */
#define GO1MPC_NLP_ROW_T_INT 367         /* _t_int: planner tick of the last walking call */
#define GO1MPC_NLP_ROW_MPC_STOP 368      /* mpc_stop */
#define GO1MPC_NLP_ROW_RIGHT_SUPPORT 369 /* right_support */
int go1mpc_nlp_node_state_doubles(void);
int go1mpc_nlp_node_default_state(go1mpc_t *h, double *state);
int go1mpc_nlp_walkdtime_max(const go1mpc_t *h);         /* _walkdtime_max: ticks of the walk (672) */
int go1mpc_nlp_t_end_footstep(const go1mpc_t *h);        /* _t_end_footstep */
int go1mpc_nlp_node_tick_batch(go1mpc_t *h, int B, double *state_d, const int *walkdtime_d, const int *start_d, const int *cmd_d,
                               const double *rfoot_fb_d, const double *lfoot_fb_d, double *msg_d, void *stream);
/* host buffers of the same SoA shapes (synchronous: copies up, tick, copies down) */
int go1mpc_nlp_node_tick_batch_host(go1mpc_t *h, int B, double *state, const int *walkdtime, const int *start, const int *cmd,
                                    const double *rfoot_fb, const double *lfoot_fb, double *msg);

/* ---------------------------------------------------------------------------
 * The 100 Hz node of rt_mpc_qp for B robots: /MPC/Gait message in, /rtMPC/traj message out.  Replaces one pass of the
 * main loop of RT/gait_fast.cpp:505-746 with everything it calls:
 *   xget_position_interpolation (:113-372) + PRMPCClass::XGetSolution_position_mod3 (RT/FastMPC/PRMPCClass.cpp:1170-1261),
 *   PRMPCClass::Foot_trajectory_solve_mod2 (:1756-2195), XGetSolution_Foot_rotation (:2255-2380), the 2 x nh reference
 *   windows (gait_fast.cpp:568-616, with their row-0-twice quirk), PRMPCClass::body_theta_mpc, the outgoing 100 slots
 *   (:633-729; slot 86 -- wall time -- is 0).
 * state_d  [go1mpc_rt_node_state_doubles(nh)][B] SoA, in/out: the node's and the class's members
 *          (go1mpc_rt_node_default_state gives one robot's initial column)
 * msg_d    [100][B] SoA: the latest message of the 40 Hz planner per robot (slots as NLPRTControlClass.cpp:288-392)
 * ctrl_d   [B] ints (control_gait(0) > 0) or NULL = on;  bodyangle_state_d [4][B] or NULL = 0
 * body_in_d  [B][go1mpc_body_in_stride(nh)] workspace; body_out_d [B][go1mpc_body_out_stride(nh)] the body MPC's records
 *          (in/out: its state lives there; zero-initialise);  body_diag_d as go1mpc_body_mpc_step_batch or NULL
 * out100_d [100][B] SoA;  active_d [B] ints or NULL: 1 where the fast tick ran (message slot 99 > 0)
 * Three launches on `stream`.  Parity: tests/test_gpu_rt_node.py (cfg1 lock-step replay of the unmodified classes).
 * ------------------------------------------------------------------------ */
int go1mpc_rt_node_state_doubles(int nh);
int go1mpc_rt_node_default_state(go1mpc_t *h, int nh, double *state);
int go1mpc_rt_node_tick_batch(go1mpc_t *h, int nh, int B, double *state_d, const double *msg_d, const int *ctrl_d,
                              const double *bodyangle_state_d, double *body_in_d, double *body_out_d, int *body_diag_d,
                              double *out100_d, int *active_d, void *stream);
/* The same tick with the node's other two topics in their wire format (gait_fast.cpp:92-110, 519-527): ctl_msg_d [25][B] is
 * /control2rtmpc/state -- slot 0 the control flag control_gait(0), slots 10, 11, 13, 14 the measured body angles / rates the
 * MPC is fed (bodyangle_state) --, rt2nrt_msg_d [25][B] (or NULL) receives /rt2nrt/state: slot 0 = t_int, slots 1..24 = the
 * control message's. */
int go1mpc_rt_node_tick_msgs_batch(go1mpc_t *h, int nh, int B, double *state_d, const double *gait_msg_d, const double *ctl_msg_d,
                                   double *body_in_d, double *body_out_d, int *body_diag_d, double *traj_msg_d,
                                   double *rt2nrt_msg_d, void *stream);
/* host buffers (synchronous); body_out [B][go1mpc_body_out_stride(nh)] is the body MPC's state, in/out */
int go1mpc_rt_node_tick_batch_host(go1mpc_t *h, int nh, int B, double *state, const double *msg, const int *ctrl,
                                   const double *bodyangle_state, double *body_out, double *out100);

/* Per-kernel timing of the three-launch body tick (measurement aid for bench.py's roofline): while enabled, each
 * go1mpc_body_mpc_step_batch call on that path records CUDA events around its setup / solve / merge launches;
 * go1mpc_body_phase_ms waits for the last such call and returns the three durations in milliseconds.  Do not enable
 * inside a graph capture. */
int go1mpc_body_phase_timing(go1mpc_t *h, int enable);
int go1mpc_body_phase_ms(go1mpc_t *h, float *ms3);

/* Measured FP64 FMA throughput of the device (GFLOP/s, 2 flop per FMA) from a
 * register-resident DFMA loop: the roofline denominator bench.py reports
 * against.  Runs ~`ms` milliseconds. */
int go1mpc_measure_dfma_peak(go1mpc_t *h, int ms, double *gflops);

#ifdef __cplusplus
}
#endif
#endif /* GO1MPC_H */
