// C-callable wrapper around the UNMODIFIED reference step-location / step-timing planner
// (NLP/src/NLP/NLPClass.h + NLPClass_sqp.cpp, NLP = unitree_ros/mosek_nlp_kmp), compiled from
// /root/reference against oracle/eigen_shim and the Mosek / Armadillo / KMP stand-ins.
// Test infrastructure only: pins oracle/step_timing.c.
#define private public
#define protected public
#include <NLP/NLPClass.h>
#include <NLPRTControl/NLPRTControlClass.h>
#undef private
#undef protected
#include <cstdlib>
#include <new>

// KMP is dead code on this path: no-op definitions for the linker.
kmp::kmp() {}
void kmp::kmp_initialize(mat&, int, int, int, double, double) {}
int kmp::kernel_extend(vec, vec, mat&) { return 0; }
int kmp::kmp_estimateMatrix() { return 0; }
int kmp::kmp_prediction(vec, vec&) { return 0; }
int kmp::kmp_insertPoint(vec) { return 0; }

extern "C" {

// same set-up as NLPRTControlClass::NLPRTControlClass (NLPRTControl/NLPRTControlClass.cpp:25-73)
void* ref_nlp_new(double steplength, double stepwidth, double stepheight) {
  NLPClass* p = new NLPClass();
  p->_method_flag = 2;
  p->_pvFlag_kmp = 1;
  p->_robot_name = "go1";
  p->_robot_mass = 12;
  p->_lift_height = 0.03;
  p->RobotPara_totalmass = 12;
  p->RobotPara_HALF_HIP_WIDTH = 0.12675;
  p->RobotPara_dt = 0.025;
  p->RobotPara_Tstep = 0.7;
  p->RobotPara_Z_C = 0.309458;
  p->RobotPara_g = 9.8;
  p->RobotPara_FOOT_WIDTH = 0.03;
  p->FootStepInputs(stepwidth, steplength, stepheight);
  p->Initialize();
  return p;
}
void ref_nlp_free(void* h) { delete static_cast<NLPClass*>(h); }

// constants the oracle's config is filled from (order = orc_step_cfg doubles)
void ref_nlp_consts(void* h, double* c) {
  NLPClass* p = static_cast<NLPClass*>(h);
  int k = 0;
  c[k++] = _dt; c[k++] = p->_Wn; c[k++] = p->_ggg(0);
  c[k++] = p->_t_min; c[k++] = p->_t_max;
  c[k++] = p->_footx_max; c[k++] = p->_footx_min;
  c[k++] = p->_footx_vmax; c[k++] = p->_footx_vmin; c[k++] = p->_footy_vmax; c[k++] = p->_footy_vmin;
  c[k++] = p->_comax_max; c[k++] = p->_comax_min; c[k++] = p->_comay_max; c[k++] = p->_comay_min;
  c[k++] = p->_aax; c[k++] = p->_aay; c[k++] = p->_aaxv; c[k++] = p->_aayv;
  c[k++] = p->_bbx; c[k++] = p->_bby; c[k++] = p->_rr1; c[k++] = p->_rr2;
  c[k++] = p->RobotPara_HALF_HIP_WIDTH; c[k++] = p->RobotPara_FOOT_WIDTH;
  c[k++] = 0; c[k++] = 0; c[k++] = 0; c[k++] = 0;   // lamda (hard-coded 0 in the reference)
  c[k++] = p->_hcom;
}

// state the tick reads: tables (7 x 27) | feed at i-1 (6) | Vari_ini.col(i-1) (4) | endref (2) | _bjx1
void ref_nlp_get_state(void* h, int i, double* s) {
  NLPClass* p = static_cast<NLPClass*>(h);
  int k = 0;
  for (int j = 0; j < 27; j++) s[k++] = p->_ts(j);
  for (int j = 0; j < 27; j++) s[k++] = p->_tx(j);
  for (int j = 0; j < 27; j++) s[k++] = p->_footx_ref(j);
  for (int j = 0; j < 27; j++) s[k++] = p->_footy_ref(j);
  for (int j = 0; j < 27; j++) s[k++] = p->_footz_ref(j);
  for (int j = 0; j < 27; j++) s[k++] = p->_Lxx_ref(j);
  for (int j = 0; j < 27; j++) s[k++] = p->_Lyy_ref(j);
  s[k++] = p->_comx_feed(i - 1); s[k++] = p->_comvx_feed(i - 1); s[k++] = p->_comax_feed(i - 1);
  s[k++] = p->_comy_feed(i - 1); s[k++] = p->_comvy_feed(i - 1); s[k++] = p->_comay_feed(i - 1);
  for (int j = 0; j < 4; j++) s[k++] = p->_Vari_ini(j, i - 1);
  s[k++] = p->_comvx_endref(0); s[k++] = p->_comvy_endref(0);
  s[k++] = p->_bjx1;
}
void ref_nlp_set_state(void* h, int i, const double* s) {
  NLPClass* p = static_cast<NLPClass*>(h);
  int k = 0;
  for (int j = 0; j < 27; j++) p->_ts(j) = s[k++];
  for (int j = 0; j < 27; j++) p->_tx(j) = s[k++];
  for (int j = 0; j < 27; j++) p->_footx_ref(j) = s[k++];
  for (int j = 0; j < 27; j++) p->_footy_ref(j) = s[k++];
  for (int j = 0; j < 27; j++) p->_footz_ref(j) = s[k++];
  for (int j = 0; j < 27; j++) p->_Lxx_ref(j) = s[k++];
  for (int j = 0; j < 27; j++) p->_Lyy_ref(j) = s[k++];
  p->_comx_feed(i - 1) = s[k++]; p->_comvx_feed(i - 1) = s[k++]; p->_comax_feed(i - 1) = s[k++];
  p->_comy_feed(i - 1) = s[k++]; p->_comvy_feed(i - 1) = s[k++]; p->_comay_feed(i - 1) = s[k++];
  for (int j = 0; j < 4; j++) p->_Vari_ini(j, i - 1) = s[k++];
  p->_comvx_endref(0) = s[k++]; p->_comvy_endref(0) = s[k++];
  p->_bjx1 = (int)s[k++];
  p->_td = 0.2 * p->_ts;
  // members the end of the previous tick leaves behind (NLPClass_sqp.cpp:1041-1046)
  for (int j = 0; j < 27; j++) { p->_footxyz_real(0, j) = p->_footx_ref(j); p->_footxyz_real(1, j) = p->_footy_ref(j); p->_footxyz_real(2, j) = p->_footz_ref(j); }
  p->_footxyz_real(1, 0) = -p->_stepwidth(0);
}

// NLPClass::step_timing_opti_loop, NLPClass_sqp.cpp:693-1102.  est18/rfoot/lfoot as the
// reference takes them; hz receives what CoM_height_solve wrote: comz[3] comaz[3] zsc[3] comvz(i);
// ints: periond_i, k_yu, bjxx, bjx1.
void ref_nlp_step(void* h, int i, const double* est18, const double* rfoot3, const double* lfoot3, int stop,
                  double* out38, double* hz, int* ints) {
  NLPClass* p = static_cast<NLPClass*>(h);
  Eigen::Matrix<double, 18, 1> est;
  Eigen::Vector3d rf, lf;
  for (int k = 0; k < 18; k++) est(k) = est18[k];
  for (int k = 0; k < 3; k++) { rf(k) = rfoot3[k]; lf(k) = lfoot3[k]; }
  Eigen::Matrix<double, 38, 1> o = p->step_timing_opti_loop(i, est, rf, lf, 0.0, stop != 0);
  for (int k = 0; k < 38; k++) out38[k] = o(k);
  for (int q = 0; q < 3; q++) { hz[q] = p->_comz(i + q); hz[3 + q] = p->_comaz(i + q); hz[6 + q] = p->_Zsc(i + q); }
  hz[9] = p->_comvz(i);
  ints[0] = p->_periond_i; ints[1] = p->_k_yu; ints[2] = p->_bjxx; ints[3] = p->_bjx1;
}

// NLPClass::Foot_trajectory_solve_mod2, NLPClass_sqp.cpp:2039-2358 (call it right after ref_nlp_step
// with the same tick, as NLPRTControlClass::rt_nlp_gait does)
int ref_nlp_foot(void* h, int j, int stop, double* out18) {
  NLPClass* p = static_cast<NLPClass*>(h);
  Eigen::Matrix<double, 18, 1> o = p->Foot_trajectory_solve_mod2(j, stop != 0);
  for (int k = 0; k < 18; k++) out18[k] = o(k);
  return p->right_support;
}
double ref_nlp_stepwidth0(void* h) { return static_cast<NLPClass*>(h)->_stepwidth(0); }

// The 40 Hz node's driver, UNMODIFIED: NLPRTControlClass::WalkingReactStepping (NLPRTControl/NLPRTControlClass.cpp:191-396)
// -> the 100-slot /MPC/Gait message.  The constructor reads the never-initialised member stepheightinput (:80): the object
// is built in zero-filled memory so that it reads 0 (flat ground), deterministically.
void* ref_ctl_new() {
  void* mem = calloc(1, sizeof(NLPRTControlClass));
  return new (mem) NLPRTControlClass();
}
void ref_ctl_free(void* h) { static_cast<NLPRTControlClass*>(h)->~NLPRTControlClass(); free(h); }
void ref_ctl_step(void* h, int count, int start, const double* est18, const double* rfoot3, const double* lfoot3, double* out100) {
  NLPRTControlClass* p = static_cast<NLPRTControlClass*>(h);
  Eigen::Matrix<double, 18, 1> est;
  Eigen::Vector3d rf, lf;
  for (int k = 0; k < 18; k++) est(k) = est18[k];
  for (int k = 0; k < 3; k++) { rf(k) = rfoot3[k]; lf(k) = lfoot3[k]; }
  Eigen::Matrix<double, 100, 1> o = p->WalkingReactStepping(count, start != 0, est, rf, lf);
  for (int k = 0; k < 100; k++) out100[k] = o(k);
}
void ref_ctl_start(void* h) { static_cast<NLPRTControlClass*>(h)->StartWalking(); }
void ref_ctl_stop(void* h) { static_cast<NLPRTControlClass*>(h)->StopWalking(); }
int ref_ctl_walkdtime_max(void* h) { return static_cast<NLPRTControlClass*>(h)->_walkdtime_max; }

}  // extern "C"
