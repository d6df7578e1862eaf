// C-callable wrapper around the UNMODIFIED reference body-inclination MPC
// (RT/src/FastMPC/PRMPCClass.{h,cpp}), compiled from /root/reference against
// oracle/eigen_shim and the Armadillo / KMP stand-ins.  Test infrastructure
// only: pins oracle/body_mpc.c at the reference's compile-time horizon _nh = 4.
#define private public
#define protected public
#include <FastMPC/PRMPCClass.h>
#undef private
#undef protected

// The KMP swing-leg generator is dead code on this path (every live call is
// commented out in the reference); Initialize() only calls kmp_initialize on a
// data file that does not exist.  No-op definitions satisfy the linker.
kmp::kmp() {}
void kmp::kmp_initialize(mat&, int, int, int, double, double) {}
int kmp::kernel_extend(vec, vec, mat&) { return 0; }
int kmp::kmp_estimateMatrix() { return 0; }
int kmp::kmp_prediction(vec, vec&) { return 0; }
int kmp::kmp_insertPoint(vec) { return 0; }

extern "C" {

int ref_body_nh() { return _nh; }

void* ref_body_new() {
  PRMPCClass* p = new PRMPCClass();
  p->Initialize();
  // members the reference never initialises (UB there); zero here and in the oracle
  p->_zmpx_real.setZero(); p->_zmpy_real.setZero();
  p->_thetax_real.setZero(); p->_thetay_real.setZero();
  return p;
}
void ref_body_free(void* h) { delete static_cast<PRMPCClass*>(h); }

// state = (thetaxk0, thetaxk1, thetayk0, thetayk1)
void ref_body_get_state(void* h, double* tx27, double* state4, double* vini) {
  PRMPCClass* p = static_cast<PRMPCClass*>(h);
  for (int i = 0; i < 27; i++) tx27[i] = p->_tx(i);
  state4[0] = p->_thetaxk(0); state4[1] = p->_thetaxk(1);
  state4[2] = p->_thetayk(0); state4[3] = p->_thetayk(1);
  for (int i = 0; i < 2 * _nh; i++) vini[i] = p->_V_ini(i);
}
void ref_body_set_state(void* h, const double* tx27, const double* state4, const double* vini) {
  PRMPCClass* p = static_cast<PRMPCClass*>(h);
  for (int i = 0; i < 27; i++) p->_tx(i) = tx27[i];
  p->_thetaxk(0) = state4[0]; p->_thetaxk(1) = state4[1];
  p->_thetayk(0) = state4[2]; p->_thetayk(1) = state4[3];
  for (int i = 0; i < 2 * _nh; i++) p->_V_ini(i) = vini[i];
}
void ref_body_get_model(void* h, double* pps, double* pvs, double* ppu, double* pvu, double* ppu2, double* pvu2) {
  PRMPCClass* p = static_cast<PRMPCClass*>(h);
  for (int j = 0; j < 2; j++) for (int i = 0; i < _nh; i++) { pps[j * _nh + i] = p->_pps(i, j); pvs[j * _nh + i] = p->_pvs(i, j); }
  for (int j = 0; j < _nh; j++) for (int i = 0; i < _nh; i++) {
    ppu[j * _nh + i] = p->_ppu(i, j); pvu[j * _nh + i] = p->_pvu(i, j);
    ppu2[j * _nh + i] = p->_ppu_2(i, j); pvu2[j * _nh + i] = p->_pvu_2(i, j);
  }
}

// PRMPCClass::body_theta_mpc, PRMPCClass.cpp:379-714.  refs in the oracle's
// layout (signal-major: x[nh] then y[nh]); comacc_z = row 2 of the 3x5 matrix.
void ref_body_theta_mpc(void* h, int i, const double* state4, const double* zmp, const double* ang,
                        const double* rfoot, const double* lfoot, const double* comacc_z,
                        double* out14, int* qp_solution) {
  PRMPCClass* p = static_cast<PRMPCClass*>(h);
  Eigen::Matrix<double, 4, 1> st;
  Eigen::Matrix<double, 2, 5> Z, A, R, L;
  Eigen::Matrix<double, 3, 5> C;
  Eigen::Matrix<double, 9, 1> N;
  Z.setZero(); A.setZero(); R.setZero(); L.setZero(); C.setZero(); N.setZero();
  for (int k = 0; k < 4; k++) st(k) = state4[k];
  for (int k = 0; k < _nh; k++) {
    Z(0, k) = zmp[k]; Z(1, k) = zmp[_nh + k];
    A(0, k) = ang[k]; A(1, k) = ang[_nh + k];
    R(0, k) = rfoot[k]; R(1, k) = rfoot[_nh + k];
    L(0, k) = lfoot[k]; L(1, k) = lfoot[_nh + k];
    C(2, k) = comacc_z[k];
  }
  Eigen::Matrix<double, 14, 1> o = p->body_theta_mpc(i, st, Z, A, R, L, C, N);
  for (int k = 0; k < 14; k++) out14[k] = o(k);
  *qp_solution = p->qp_solution ? 1 : 0;
}

// XGetSolution_position_mod3, PRMPCClass.cpp:1170-1261 (uses _AAA_inv_mod of solve_AAA_inv_mod1 :1344-1361, built by
// Initialize).  out21 = the Vec21; inv16 (may be null) = _AAA_inv_mod row-major; returns _t_end_footstep.
int ref_body_position_mod3(void* h, int walktime, double dt_sample, const double* in1, const double* in2, const double* ref,
                           const double* ref2, double* out21, double* inv16) {
  PRMPCClass* p = static_cast<PRMPCClass*>(h);
  Eigen::Vector3d a, b, c, d;
  for (int k = 0; k < 3; k++) { a(k) = in1[k]; b(k) = in2[k]; c(k) = ref[k]; d(k) = ref2[k]; }
  Eigen::Matrix<double, 21, 1> o = p->XGetSolution_position_mod3(walktime, dt_sample, a, b, c, d);
  for (int k = 0; k < 21; k++) out21[k] = o(k);
  if (inv16) for (int r = 0; r < 4; r++) for (int k = 0; k < 4; k++) inv16[4 * r + k] = p->_AAA_inv_mod(r, k);
  return (int)p->_t_end_footstep;
}

// XGetSolution_Foot_rotation, PRMPCClass.cpp:2255-2380: tables it reads, a setter for the step lengths, the call.
void ref_body_foot_tables(void* h, double* tx27, double* ts27, double* td27, double* footx27, double* scal4) {
  PRMPCClass* p = static_cast<PRMPCClass*>(h);
  for (int k = 0; k < _footstepsnumber; k++) { tx27[k] = p->_tx(k); ts27[k] = p->_ts(k); td27[k] = p->_td(k); footx27[k] = p->_footxyz_real(0, k); }
  scal4[0] = p->_footx_max; scal4[1] = _dt_mpc; scal4[2] = p->_t_end_footstep; scal4[3] = p->_bjx1;
}
void ref_body_set_footx(void* h, const double* footx27) {
  PRMPCClass* p = static_cast<PRMPCClass*>(h);
  for (int k = 0; k < _footstepsnumber; k++) p->_footxyz_real(0, k) = footx27[k];
}
void ref_body_foot_rotation(void* h, int walktimex, double dt_sample, double* out30) {
  PRMPCClass* p = static_cast<PRMPCClass*>(h);
  Eigen::Matrix<double, 30, 1> o = p->XGetSolution_Foot_rotation(walktimex, dt_sample);
  for (int k = 0; k < 30; k++) out30[k] = o(k);
}

// Foot_trajectory_solve_mod2, PRMPCClass.cpp:1756-2195: the call, and the members it leaves behind
// (state layout of oracle/rt_foot.c at the reference's nh: ts 27 | footxyz_real 81 | lift 27 | ry | bjxx | bjx1 | six arrays of _nh + 2).
void ref_body_foot_traj(void* h, int j_indexx, int stop, const double* nrt9, double* out30) {
  PRMPCClass* p = static_cast<PRMPCClass*>(h);
  Eigen::Matrix<double, 9, 1> N;
  for (int k = 0; k < 9; k++) N(k) = nrt9[k];
  Eigen::Matrix<double, 30, 1> o = p->Foot_trajectory_solve_mod2(j_indexx, stop != 0, N);
  for (int k = 0; k < 30; k++) out30[k] = o(k);
}
void ref_body_foot_traj_state(void* h, double* s) {
  PRMPCClass* p = static_cast<PRMPCClass*>(h);
  int k = 0;
  for (int j = 0; j < 27; j++) s[k++] = p->_ts(j);
  for (int r = 0; r < 3; r++) for (int j = 0; j < 27; j++) s[k++] = p->_footxyz_real(r, j);
  for (int j = 0; j < 27; j++) s[k++] = p->_lift_height_ref(j);
  s[k++] = p->_ry_left_right; s[k++] = p->_bjxx; s[k++] = p->_bjx1;
  const int W = _nh + 2;
  for (int j = 0; j < W; j++) s[k++] = p->_Rfootx(j);
  for (int j = 0; j < W; j++) s[k++] = p->_Rfooty(j);
  for (int j = 0; j < W; j++) s[k++] = p->_Rfootz(j);
  for (int j = 0; j < W; j++) s[k++] = p->_Lfootx(j);
  for (int j = 0; j < W; j++) s[k++] = p->_Lfooty(j);
  for (int j = 0; j < W; j++) s[k++] = p->_Lfootz(j);
}

// Hooks of oracle/rt_glue.c onto the UNMODIFIED class (ctx = the PRMPCClass object): the glue of gait_fast.cpp is restated
// once, in C, and drives either these or the oracle restatements.  Windows are signal-major [2][nh] with nh = _nh.
void ref_hook_mod3(void* h, int nh, int walktime, double dts, const double* a, const double* b, const double* r, const double* r2, double* out) {
  double o21[21];
  ref_body_position_mod3(h, walktime, dts, a, b, r, r2, o21, nullptr);
  for (int k = 0; k < 9 + 3 * (nh - 1); k++) out[k] = o21[k];
}
void ref_hook_foot(void* h, int nh, int j, int stop, const double* nrt9, double* out) {
  double o30[30];
  ref_body_foot_traj(h, j, stop, nrt9, o30);
  for (int k = 0; k < 6 * (nh + 1); k++) out[k] = o30[k];
}
void ref_hook_rot(void* h, int nh, int j, double dts, double* out) {
  double o30[30];
  ref_body_foot_rotation(h, j, dts, o30);
  for (int k = 0; k < 6 * nh; k++) out[k] = o30[k];
}
void ref_hook_body(void* h, int nh, int i, const double* bs, const double* zmp, const double* ang, const double* rf, const double* lf,
                   const double* acc, double* out14) {
  int ok;
  (void)nh;
  ref_body_theta_mpc(h, i, bs, zmp, ang, rf, lf, acc, out14, &ok);
}
double ref_hook_tx_total(void* h) { return (double)(int)static_cast<PRMPCClass*>(h)->_tx_total; }

}  // extern "C"
