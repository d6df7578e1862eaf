// go1mpc.hpp -- headless C++ mirror of the reference MPC classes' call surface over the C ABI
// (include/go1mpc.h).  No ROS, Gazebo, Eigen, Boost, Armadillo or Mosek: fixed-size vectors are
// the small Vec/Mat templates below (column-major, operator()(i[,j]) like Eigen), so call sites
// written against the reference keep their shape:
//
//   reference class (file)                                   mirror
//   QPBaseClass      RT/src/QP/QPBaseClass.{h,cpp}           go1host::QPBase
//   PRMPCClass       RT/src/FastMPC/PRMPCClass.{h,cpp}       go1host::BodyInclinationMPC   (body_theta_mpc)
//   NLPClass         NLP/src/NLP/NLPClass.h, NLPClass_sqp.cpp go1host::StepTimingMPC       (step_timing_opti_loop)
//   Kinematicclass   GO1/src/kinematics/Kinematics.{h,cpp}   go1host::LegKinematics        (+ Jacobian_kin)
//
// Same names, argument order and meaning, same error behaviour (no exceptions from the solve
// calls: a failed QP shows as solveQP() == false / qp_solution == false and the reference's
// fallbacks apply).  Every call is a batch of one through the *_host entry points; the
// *_batch members expose the batched device path for callers that own many instances.
// A missing GPU is a constructor-time std::runtime_error: there is no CPU fallback.
#pragma once
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/go1mpc.h"

namespace go1host {

template <int N>
struct Vec {
  double v[N];
  Vec() { std::memset(v, 0, sizeof v); }
  double& operator()(int i) { return v[i]; }
  double operator()(int i) const { return v[i]; }
  double& operator()(int i, int) { return v[i]; }
  double operator()(int i, int) const { return v[i]; }
  void setZero() { std::memset(v, 0, sizeof v); }
  static constexpr int size() { return N; }
};
template <int R, int C>
struct Mat {   // column-major, like Eigen's default
  double m[R * C];
  Mat() { std::memset(m, 0, sizeof m); }
  double& operator()(int r, int c) { return m[c * R + r]; }
  double operator()(int r, int c) const { return m[c * R + r]; }
  void setZero() { std::memset(m, 0, sizeof m); }
};
// run-time sized reference window: rows x cols, column-major (the reference's fixed 2x5 / 3x5
// matrices widened to 2 x nh / 3 x nh)
struct MatX {
  int rows = 0, cols = 0;
  std::vector<double> m;
  MatX() {}
  MatX(int r, int c) : rows(r), cols(c), m((size_t)r * c, 0.0) {}
  double& operator()(int r, int c) { return m[(size_t)c * rows + r]; }
  double operator()(int r, int c) const { return m[(size_t)c * rows + r]; }
};

// shared CUDA handle (one stream); thread-compatible, not thread-safe -- like the reference objects
class Context {
 public:
  explicit Context(int device = -1, const Go1MpcConfig* cfg = nullptr) {
    Go1MpcConfig c;
    if (cfg) c = *cfg; else go1mpc_config_default(&c);
    int rc = go1mpc_create(&c, device, &h_);
    if (rc != GO1MPC_OK) throw std::runtime_error("go1mpc_create failed (" + std::to_string(rc) + "): no usable CUDA device; there is no CPU fallback");
  }
  ~Context() { go1mpc_destroy(h_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  go1mpc_t* get() const { return h_; }
  static std::shared_ptr<Context> shared(int device = -1) {
    static std::weak_ptr<Context> w;
    auto s = w.lock();
    if (!s) { s = std::make_shared<Context>(device); w = s; }
    return s;
  }
 private:
  go1mpc_t* h_ = nullptr;
};

// ---------------------------------------------------------------------------------------------
// QPBaseClass: problem container + solveQP().  min 0.5 x'Gx + g0'x, CE'x + ce0 = 0, CI'x + ci0 >= 0
// (RT/src/QP/QPBaseClass.h:60-68, QPBaseClass.cpp:102-153).  Matrices column-major.
class QPBase {
 public:
  explicit QPBase(const std::string& qpSolverName = "EiQuadProg", std::shared_ptr<Context> ctx = nullptr)
      : ctx_(ctx ? ctx : Context::shared()) {
    if (qpSolverName != "EiQuadProg") throw std::runtime_error("only the EiQuadProg backend exists (as in the reference)");
  }
  void resizeQP(int nVars, int nEqCon, int nIneqCon) {
    _nVars = nVars; _nEqCon = nEqCon; _nIneqCon = nIneqCon;
    _G.assign((size_t)nVars * nVars, 0.0); _g0.assign(nVars, 0.0);
    _CE.assign((size_t)nVars * nEqCon, 0.0); _ce0.assign(nEqCon, 0.0);
    _CI.assign((size_t)nVars * nIneqCon, 0.0); _ci0.assign(nIneqCon, 0.0);
    _X.assign(nVars, 0.0); _active.assign(nIneqCon + nEqCon, 0);
  }
  // returns the reference's success flag: "no NaN in X" (the cost, +inf when infeasible / not PD, is dropped there)
  bool solveQP() {
    int it[GO1MPC_ITERS];
    int rc = go1mpc_qp_solve_batch_host(ctx_->get(), _nVars, _nEqCon, _nIneqCon, 1, _G.data(), _g0.data(),
                                        _nEqCon ? _CE.data() : nullptr, _nEqCon ? _ce0.data() : nullptr,
                                        _nIneqCon ? _CI.data() : nullptr, _nIneqCon ? _ci0.data() : nullptr,
                                        _X.data(), &_cost, _active.data(), &_nactive, it, &_status);
    if (rc != GO1MPC_OK) { _status = GO1MPC_QP_NAN; return false; }
    for (double x : _X) if (x != x) return false;
    return true;
  }
  double& G(int i, int j) { return _G[(size_t)j * _nVars + i]; }
  double& CE(int i, int j) { return _CE[(size_t)j * _nVars + i]; }
  double& CI(int i, int j) { return _CI[(size_t)j * _nVars + i]; }
  int _nVars = 0, _nEqCon = 0, _nIneqCon = 0;
  std::vector<double> _G, _g0, _CE, _ce0, _CI, _ci0, _X;
  // extras the reference does not expose
  double _cost = 0.0; int _status = 0, _nactive = 0; std::vector<int> _active;
 protected:
  std::shared_ptr<Context> ctx_;
};

// ---------------------------------------------------------------------------------------------
// PRMPCClass (body-inclination MPC part).  body_theta_mpc keeps the reference's signature
// (RT/src/FastMPC/PRMPCClass.h:67-70, PRMPCClass.cpp:379-714) with the 2x5 / 3x5 reference
// windows widened to 2 x nh / 3 x nh; the horizon is a constructor argument (reference: 4).
class BodyInclinationMPC {
 public:
  explicit BodyInclinationMPC(int nh = 4, std::shared_ptr<Context> ctx = nullptr)
      : _nh(nh), ctx_(ctx ? ctx : Context::shared()) { Initialize(); }
  void Initialize() {
    _thetaxk.setZero(); _thetayk.setZero(); _V_ini.assign(2 * _nh, 0.0); out14_.setZero();
    go1mpc_body_default_tx(ctx_->get(), _tx.v);
    qp_solution = true; _bjx1 = _bjx2 = 0;
  }
  Vec<14> body_theta_mpc(int i, const Vec<4>& bodyangle_state, const MatX& zmp_mpc_ref, const MatX& bodyangle_mpc_ref,
                         const MatX& rfoot_mpc_ref, const MatX& lfoot_mpc_ref, const MatX& comacc_mpc_ref,
                         const Vec<9>& /*Nrtfoorpr_gen: unused on this path, as in the reference*/) {
    const int nh = _nh, is = go1mpc_body_in_stride(nh), os = go1mpc_body_out_stride(nh), ds = go1mpc_body_diag_stride(nh);
    std::vector<double> in(is, 0.0), out(os, 0.0);
    std::vector<int> diag(ds, 0);
    std::memcpy(in.data(), _tx.v, sizeof(double) * 27);
    in[27] = (double)i;
    in[28] = _thetaxk(0); in[29] = _thetaxk(1); in[30] = _thetayk(0); in[31] = _thetayk(1);
    for (int k = 0; k < 4; k++) in[32 + k] = bodyangle_state(k);
    for (int k = 0; k < 2 * nh; k++) in[36 + k] = _V_ini[k];
    double* r = in.data() + 36 + 2 * nh;
    for (int k = 0; k < nh; k++) {
      r[k] = zmp_mpc_ref(0, k); r[nh + k] = zmp_mpc_ref(1, k);
      r[2 * nh + k] = bodyangle_mpc_ref(0, k); r[3 * nh + k] = bodyangle_mpc_ref(1, k);
      r[4 * nh + k] = rfoot_mpc_ref(0, k); r[5 * nh + k] = rfoot_mpc_ref(1, k);
      r[6 * nh + k] = lfoot_mpc_ref(0, k); r[7 * nh + k] = lfoot_mpc_ref(1, k);
      r[8 * nh + k] = comacc_mpc_ref(2, k);
    }
    for (int k = 0; k < 14; k++) out[k] = out14_(k);      // gated ticks return the stale members
    int rc = go1mpc_body_mpc_step_batch_host(ctx_->get(), nh, 1, in.data(), out.data(), diag.data());
    if (rc != GO1MPC_OK) throw std::runtime_error(std::string("body_theta_mpc: ") + go1mpc_last_error(ctx_->get()));
    for (int k = 0; k < 14; k++) out14_(k) = out[k];
    _thetaxk(0) = out[14]; _thetaxk(1) = out[15]; _thetayk(0) = out[16]; _thetayk(1) = out[17];
    for (int k = 0; k < 2 * nh; k++) _V_ini[k] = out[18 + k];
    _status = diag[0]; qp_solution = !(diag[0] == GO1MPC_QP_NAN);
    _bjx1 = diag[6]; _bjx2 = diag[7];
    _active.assign(diag.begin() + 10, diag.begin() + 10 + (diag[1] > 0 ? diag[1] : 0));
    return out14_;
  }
  int _nh;
  Vec<2> _thetaxk, _thetayk;
  std::vector<double> _V_ini;
  Vec<27> _tx;
  bool qp_solution = true;
  int _bjx1 = 0, _bjx2 = 0, _status = 0;
  std::vector<int> _active;
 private:
  Vec<14> out14_;
  std::shared_ptr<Context> ctx_;
};

// ---------------------------------------------------------------------------------------------
// NLPClass (step-location / step-timing planner).  FootStepInputs / Initialize /
// step_timing_opti_loop keep the reference's signatures (NLP/src/NLP/NLPClass.h:64-94).
class StepTimingMPC {
 public:
  explicit StepTimingMPC(std::shared_ptr<Context> ctx = nullptr) : ctx_(ctx ? ctx : Context::shared()) {
    FootStepInputs(0.2535, 0.075, 0.0); Initialize();
  }
  void FootStepInputs(double stepwidth, double steplengthx, double stepheight) { sw_ = stepwidth; sl_ = steplengthx; sh_ = stepheight; }
  void Initialize() {
    state.assign(GO1MPC_STEP_STATE_DOUBLES, 0.0);
    go1mpc_step_default_state(ctx_->get(), sl_, sw_, sh_, 0.7, state.data());
    foot_state.assign(GO1MPC_FOOT_STATE_DOUBLES, 0.0);
    go1mpc_foot_default_state(ctx_->get(), foot_state.data());
    last_out38_.assign(GO1MPC_STEP_OUT_DOUBLES, 0.0);
    _periond_i = _k_yu = _bjxx = _bjx1 = 0; right_support = 2; lift_zero_from = GO1MPC_FOOTSTEPS;
  }
  Vec<38> step_timing_opti_loop(int i, const Vec<18>& estimated_state, const Vec<3>& _Rfoot_location_feedback,
                                const Vec<3>& _Lfoot_location_feedback, double /*lamda*/, bool /*_stopwalking*/) {
    double in[GO1MPC_STEP_IN_DOUBLES] = {0}, out[GO1MPC_STEP_OUT_DOUBLES];
    int diag[GO1MPC_STEP_DIAG_INTS];
    for (int k = 0; k < 6; k++) in[k] = estimated_state(k);
    in[6] = _Rfoot_location_feedback(0); in[7] = _Rfoot_location_feedback(1);
    in[8] = _Lfoot_location_feedback(0); in[9] = _Lfoot_location_feedback(1);
    int rc = go1mpc_step_timing_step_batch_host(ctx_->get(), 3, 1, &i, state.data(), in, out, diag);
    if (rc != GO1MPC_OK) throw std::runtime_error(std::string("step_timing_opti_loop: ") + go1mpc_last_error(ctx_->get()));
    _periond_i = diag[0]; _k_yu = diag[1]; _bjxx = diag[2]; _bjx1 = diag[3];
    for (int q = 0; q < 5; q++) qp_status[q] = diag[5 + 11 * q];
    Vec<38> r;
    for (int k = 0; k < 38; k++) { r(k) = out[k]; last_out38_[k] = out[k]; }
    return r;
  }
  // NLPClass::Foot_trajectory_solve_mod2 (NLP/src/NLP/NLPClass_sqp.cpp:2039-2358): call it right after
  // step_timing_opti_loop with the same index, as NLPRTControlClass::rt_nlp_gait does.
  Vec<18> Foot_trajectory_solve_mod2(int j_index, bool _stopwalking) {
    double out[GO1MPC_FOOT_OUT_DOUBLES];
    const double stop = _stopwalking ? 1.0 : 0.0;
    int rc = go1mpc_foot_trajectory_stop_batch_host(ctx_->get(), 1, &j_index, state.data(), last_out38_.data(), foot_state.data(), out,
                                                    &right_support, &lift_zero_from, &stop);
    if (rc != GO1MPC_OK) throw std::runtime_error(std::string("Foot_trajectory_solve_mod2: ") + go1mpc_last_error(ctx_->get()));
    Vec<18> r;
    for (int k = 0; k < 18; k++) r(k) = out[k];
    return r;
  }
  std::vector<double> state;          // the 202-double planner state (layout: go1mpc.h)
  std::vector<double> foot_state;     // the 32-double swing-foot window
  double lift_zero_from = GO1MPC_FOOTSTEPS;   // first step whose _lift_height_ref a stop has zeroed (:2043-2048)
  int right_support = 2;
  int _periond_i, _k_yu, _bjxx, _bjx1, qp_status[5];
 private:
  double sw_ = 0.2535, sl_ = 0.075, sh_ = 0.0;
  std::vector<double> last_out38_;
  std::shared_ptr<Context> ctx_;
};

// ---------------------------------------------------------------------------------------------
// NLPRTControlClass (NLP/src/NLPRTControl/NLPRTControlClass.h): the 40 Hz planner node.  WalkingReactStepping keeps the
// reference's signature and returns the 100-slot /MPC/Gait vector; StartWalking / StopWalking act on the same flags.
class GaitPlannerNode {
 public:
  explicit GaitPlannerNode(std::shared_ptr<Context> ctx = nullptr) : ctx_(ctx ? ctx : Context::shared()) {
    state.assign(go1mpc_nlp_node_state_doubles(), 0.0);
    go1mpc_nlp_node_default_state(ctx_->get(), state.data());
    _walkdtime_max = go1mpc_nlp_walkdtime_max(ctx_->get());
  }
  // estimated_statex is accepted and ignored, as in the reference (rt_nlp_gait hands the planner its zero member, :446)
  Vec<100> WalkingReactStepping(int walkdtime, bool start_mpc, const Vec<18>& /*estimated_statex*/,
                                const Vec<3>& _Rfoot_location_feedbackx, const Vec<3>& _Lfoot_location_feedbackx) {
    const int start = start_mpc ? 1 : 0;
    Vec<100> msg;
    int rc = go1mpc_nlp_node_tick_batch_host(ctx_->get(), 1, state.data(), &walkdtime, &start, nullptr, _Rfoot_location_feedbackx.v,
                                             _Lfoot_location_feedbackx.v, msg.v);
    if (rc != GO1MPC_OK) throw std::runtime_error(std::string("WalkingReactStepping: ") + go1mpc_last_error(ctx_->get()));
    right_support = (int)state[GO1MPC_NLP_ROW_RIGHT_SUPPORT]; mpc_stop = (int)state[GO1MPC_NLP_ROW_MPC_STOP];
    _t_int = (int)state[GO1MPC_NLP_ROW_T_INT];
    return msg;
  }
  void StartWalking() {       // :400-414
    if (state[GO1MPC_NLP_ROW_STOP] != 0.0) state[GO1MPC_NLP_ROW_START_AGAIN] = 1.0;
    state[GO1MPC_NLP_ROW_STOP] = 0.0;
  }
  void StopWalking() {        // :416-432
    if (!(state[GO1MPC_NLP_ROW_T_INT] < 10)) state[GO1MPC_NLP_ROW_STOP] = 1.0;
  }
  std::vector<double> state;  // the node's members, planner state and ZMP history (layout: go1mpc.h)
  int right_support = 2, mpc_stop = 0, _t_int = 0, _walkdtime_max = 0;
 private:
  std::shared_ptr<Context> ctx_;
};

// ---------------------------------------------------------------------------------------------
// The 100 Hz node of rt_mpc_qp (RT/src/gait_fast.cpp:505-746): one pass of its main loop -- the latest /MPC/Gait message in,
// the /rtMPC/traj message out -- with the PRMPCClass object and the loop's variables as members.
class FastGaitNode {
 public:
  explicit FastGaitNode(int nh = 4, std::shared_ptr<Context> ctx = nullptr) : _nh(nh), ctx_(ctx ? ctx : Context::shared()) {
    state.assign(go1mpc_rt_node_state_doubles(nh), 0.0);
    go1mpc_rt_node_default_state(ctx_->get(), nh, state.data());
    body_state.assign(go1mpc_body_out_stride(nh), 0.0);
  }
  Vec<100> tick(const Vec<100>& mpc_gait_msg, bool control_gait_on, const Vec<4>& bodyangle_state) {
    const int ctrl = control_gait_on ? 1 : 0;
    Vec<100> out;
    int rc = go1mpc_rt_node_tick_batch_host(ctx_->get(), _nh, 1, state.data(), mpc_gait_msg.v, &ctrl, bodyangle_state.v,
                                            body_state.data(), out.v);
    if (rc != GO1MPC_OK) throw std::runtime_error(std::string("FastGaitNode::tick: ") + go1mpc_last_error(ctx_->get()));
    return out;
  }
  int _nh;
  std::vector<double> state;        // the loop's variables and the PRMPCClass members around the body MPC
  std::vector<double> body_state;   // the body MPC's output record = its state
 private:
  std::shared_ptr<Context> ctx_;
};

// ---------------------------------------------------------------------------------------------
// Kinematicclass (GO1/src/kinematics/Kinematics.h:49-60): same four methods; the Jacobian is
// left in the public member Jacobian_kin after each call, as go1_servo reads it (servo.cpp:734-741).
class LegKinematics {
 public:
  explicit LegKinematics(std::shared_ptr<Context> ctx = nullptr) : ctx_(ctx ? ctx : Context::shared()) {}
  Vec<3> Forward_kinematics(const Vec<3>& q_joint, int feet_flag) { return fk(nullptr, nullptr, q_joint, feet_flag); }
  Vec<3> Forward_kinematics_g(const Vec<3>& body_P, const Vec<3>& body_R, const Vec<3>& q_joint, int feet_flag) {
    return fk(body_P.v, body_R.v, q_joint, feet_flag);
  }
  Vec<3> Inverse_kinematics(const Vec<3>& pos_des, const Vec<3>& q_ini, int feet_flag) { return ik(nullptr, nullptr, pos_des, q_ini, feet_flag); }
  Vec<3> Inverse_kinematics_g(const Vec<3>& body_P, const Vec<3>& body_R, const Vec<3>& pos_des, const Vec<3>& q_ini, int feet_flag) {
    return ik(body_P.v, body_R.v, pos_des, q_ini, feet_flag);
  }
  Mat<3, 3> Jacobian_kin;
  int ik_updates = 0;
 private:
  Vec<3> fk(const double* bp, const double* br, const Vec<3>& q, int leg) {
    Vec<3> pos; double J[9];
    int rc = go1mpc_leg_fk_batch_host(ctx_->get(), 1, q.v, &leg, bp, br, pos.v, J);
    if (rc != GO1MPC_OK) throw std::runtime_error(std::string("Forward_kinematics: ") + go1mpc_last_error(ctx_->get()));
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) Jacobian_kin(r, c) = J[3 * r + c];
    return pos;
  }
  Vec<3> ik(const double* bp, const double* br, const Vec<3>& pdes, const Vec<3>& qini, int leg) {
    Vec<3> q; double J[9];
    int rc = go1mpc_leg_ik_batch_host(ctx_->get(), 1, pdes.v, qini.v, &leg, bp, br, q.v, J, &ik_updates);
    if (rc != GO1MPC_OK) throw std::runtime_error(std::string("Inverse_kinematics: ") + go1mpc_last_error(ctx_->get()));
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) Jacobian_kin(r, c) = J[3 * r + c];
    return q;
  }
  std::shared_ptr<Context> ctx_;
};

// ---------------------------------------------------------------------------------------------
// Dynamiccclass (GO1/src/whole_body_dynamics/dynmics_compute.h:31-91): the GRF distribution of go1_servo's 1 kHz
// loop -- force_distribution (:141-261 of the .cpp), force_opt (:265-373) and compute_joint_torques (:109-138) with
// the members the servo reads between the calls (F_leg_ref, F_leg_guess, grf_opt, qp_solution).
class GrfDistributor {
 public:
  explicit GrfDistributor(std::shared_ptr<Context> ctx = nullptr) : ctx_(ctx ? ctx : Context::shared()) {}
  Mat<3, 4> F_leg_ref;      // columns FR, FL, RR, RL
  Vec<12> F_leg_guess, grf_opt;
  bool qp_solution = false;
  double cost = 0.0;

  void force_distribution(const Vec<3>& com_des, const Vec<12>& leg_des, const Vec<6>& F_force_des, int mode, double y_coefficient,
                          const double rfoot_des[3], const double lfoot_des[3]) {
    if (mode == 101 || mode == 102) {      // other gait modes leave F_leg_ref as it is (:154-250)
      double out[12];
      int rc = go1mpc_grf_force_distribution_batch_host(ctx_->get(), 1, mode, y_coefficient, com_des.v, leg_des.v, F_force_des.v,
                                                        rfoot_des, lfoot_des, out);
      if (rc != GO1MPC_OK) throw std::runtime_error(std::string("force_distribution: ") + go1mpc_last_error(ctx_->get()));
      for (int c = 0; c < 4; c++) for (int r = 0; r < 3; r++) F_leg_ref(r, c) = out[3 * c + r];
    }
    for (int c = 0; c < 4; c++) for (int r = 0; r < 3; r++) F_leg_guess(3 * c + r) = F_leg_ref(r, c);   // :256-259
  }

  void force_opt(const Vec<3>& base_p, const Vec<3>& FR_p, const Vec<3>& FL_p, const Vec<3>& RR_p, const Vec<3>& RL_p,
                 const Vec<6>& FT_total_des, int mode, int right_support, double /*y_coefficient: unused by the reference too*/) {
    double in[48] = {0}, out[16] = {0};
    const Vec<3>* legs[4] = {&FR_p, &FL_p, &RR_p, &RL_p};
    for (int k = 0; k < 3; k++) in[k] = base_p(k);
    for (int l = 0; l < 4; l++) for (int k = 0; k < 3; k++) in[3 + 3 * l + k] = (*legs[l])(k);
    for (int k = 0; k < 6; k++) in[15 + k] = FT_total_des(k);
    for (int k = 0; k < 12; k++) { in[21 + k] = F_leg_guess(k); in[33 + k] = grf_opt(k); }
    in[45] = mode; in[46] = right_support;
    int rc = go1mpc_grf_force_opt_batch_host(ctx_->get(), 1, in, out, nullptr);
    if (rc != GO1MPC_OK) throw std::runtime_error(std::string("force_opt: ") + go1mpc_last_error(ctx_->get()));
    for (int k = 0; k < 12; k++) grf_opt(k) = out[k];
    cost = out[12]; qp_solution = out[13] != 0.0;
  }

  Vec<3> compute_joint_torques(const Mat<3, 3>& Jaco, bool support_flag, const Vec<3>& p_des, const Vec<3>& p_est,
                               const Vec<3>& pv_des, const Vec<3>& pv_est, int leg_number) {
    if (leg_number < 0 || leg_number > 3) throw std::runtime_error("compute_joint_torques: leg_number 0..3");
    double jac[36] = {0}, pd[12] = {0}, pe[12] = {0}, vd[12] = {0}, ve[12] = {0}, F[12], tau[12];
    int swing[4] = {0, 0, 0, 0};
    swing[leg_number] = support_flag ? 1 : 0;      // the reference's flag is TRUE for a swing leg (:119)
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) jac[9 * leg_number + 3 * r + c] = Jaco(r, c);
    for (int k = 0; k < 3; k++) {
      pd[3 * leg_number + k] = p_des(k); pe[3 * leg_number + k] = p_est(k);
      vd[3 * leg_number + k] = pv_des(k); ve[3 * leg_number + k] = pv_est(k);
    }
    for (int c = 0; c < 4; c++) for (int r = 0; r < 3; r++) F[3 * c + r] = F_leg_ref(r, c);
    int rc = go1mpc_grf_joint_torques_batch_host(ctx_->get(), 1, jac, swing, pd, pe, vd, ve, F, 1, 1, tau);
    if (rc != GO1MPC_OK) throw std::runtime_error(std::string("compute_joint_torques: ") + go1mpc_last_error(ctx_->get()));
    Vec<3> t;
    for (int k = 0; k < 3; k++) t(k) = tau[3 * leg_number + k];
    return t;
  }

 private:
  std::shared_ptr<Context> ctx_;
};

}  // namespace go1host
