// C-callable wrapper around the UNMODIFIED reference GRF distributor
// (GO1/src/whole_body_dynamics/dynmics_compute.{h,cpp}), compiled from /root/reference against
// oracle/eigen_shim and empty ROS header stand-ins.  Test infrastructure only: pins oracle/grf.c.
#define private public
#define protected public
#include <whole_body_dynamics/dynmics_compute.h>
#undef private
#undef protected

extern "C" {

void* ref_dyn_new() { return new Dynamiccclass(); }
void ref_dyn_free(void* h) { delete static_cast<Dynamiccclass*>(h); }

// force_distribution, dynmics_compute.cpp:141-261 -> F_leg_ref (3x4 column-major)
void ref_dyn_force_distribution(void* h, const double* com_des, const double* leg_des, const double* F6, int mode, double yc,
                                const double* rfoot, const double* lfoot, double* F_leg_ref) {
  Dynamiccclass* d = static_cast<Dynamiccclass*>(h);
  Eigen::Matrix<double, 3, 1> c; Eigen::Matrix<double, 12, 1> l; Eigen::Matrix<double, 6, 1> F;
  for (int k = 0; k < 3; k++) c(k) = com_des[k];
  for (int k = 0; k < 12; k++) l(k) = leg_des[k];
  for (int k = 0; k < 6; k++) F(k) = F6[k];
  double rf[3] = {rfoot[0], rfoot[1], rfoot[2]}, lf[3] = {lfoot[0], lfoot[1], lfoot[2]};
  d->F_leg_ref.setZero();
  d->force_distribution(c, l, F, mode, yc, rf, lf);
  for (int col = 0; col < 4; col++) for (int r = 0; r < 3; r++) F_leg_ref[col * 3 + r] = d->F_leg_ref(r, col);
}

// force_opt + solve_grf_opt, :265-427.  grf in/out = the member grf_opt; returns qp_solution.
int ref_dyn_force_opt(void* h, const double* base_p, const double* leg_p, const double* FT6, const double* F_leg_guess,
                      int mode, int right_support, double yc, double* grf) {
  Dynamiccclass* d = static_cast<Dynamiccclass*>(h);
  Eigen::Matrix<double, 3, 1> b, p[4]; Eigen::Matrix<double, 6, 1> FT;
  for (int k = 0; k < 3; k++) b(k) = base_p[k];
  for (int l = 0; l < 4; l++) for (int k = 0; k < 3; k++) p[l](k) = leg_p[3 * l + k];
  for (int k = 0; k < 6; k++) FT(k) = FT6[k];
  for (int k = 0; k < 12; k++) { d->F_leg_guess(k) = F_leg_guess[k]; d->grf_opt(k) = grf[k]; }
  d->force_opt(b, p[0], p[1], p[2], p[3], FT, mode, right_support, yc);
  for (int k = 0; k < 12; k++) grf[k] = d->grf_opt(k);
  return d->qp_solution ? 1 : 0;
}

// compute_joint_torques, :109-138.  Jaco row-major 3x3; F_ref = column leg_number of the member F_leg_ref.
void ref_dyn_joint_torques(void* h, const double* Jaco, int swing, const double* p_des, const double* p_est,
                           const double* pv_des, const double* pv_est, const double* F_ref, int leg_number, double* tau) {
  Dynamiccclass* d = static_cast<Dynamiccclass*>(h);
  Eigen::Matrix<double, 3, 3> J; Eigen::Matrix<double, 3, 1> a, b, c, e;
  for (int r = 0; r < 3; r++) for (int k = 0; k < 3; k++) J(r, k) = Jaco[3 * r + k];
  for (int k = 0; k < 3; k++) { a(k) = p_des[k]; b(k) = p_est[k]; c(k) = pv_des[k]; e(k) = pv_est[k]; d->F_leg_ref(k, leg_number) = F_ref[k]; }
  Eigen::Matrix<double, 3, 1> t = d->compute_joint_torques(J, swing != 0, a, b, c, e, leg_number);
  for (int k = 0; k < 3; k++) tau[k] = t(k);
}

}  // extern "C"
