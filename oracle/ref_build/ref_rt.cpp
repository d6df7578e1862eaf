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

}  // extern "C"
