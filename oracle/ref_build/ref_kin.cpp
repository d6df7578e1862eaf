// C-callable wrapper around the UNMODIFIED reference leg kinematics
// (GO1/src/kinematics/Kinematics.{h,cpp}), compiled from /root/reference against
// oracle/eigen_shim and empty ROS header stand-ins.  Test infrastructure only.
#include <kinematics/Kinematics.h>

extern "C" {

static void put3(const Eigen::Matrix<double, 3, 1>& v, double* o) { o[0] = v(0, 0); o[1] = v(1, 0); o[2] = v(2, 0); }
static Eigen::Matrix<double, 3, 1> get3(const double* p) {
  Eigen::Matrix<double, 3, 1> v; v(0, 0) = p[0]; v(1, 0) = p[1]; v(2, 0) = p[2]; return v;
}
static void putJ(const Kinematicclass& k, double* J) {  // row-major 3x3
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) J[i * 3 + j] = k.Jacobian_kin(i, j);
}

void ref_fk(const double* q, int leg, double* pos, double* J) {
  Kinematicclass k; put3(k.Forward_kinematics(get3(q), leg), pos); putJ(k, J);
}
void ref_fk_g(const double* bp, const double* br, const double* q, int leg, double* pos, double* J) {
  Kinematicclass k; put3(k.Forward_kinematics_g(get3(bp), get3(br), get3(q), leg), pos); putJ(k, J);
}
void ref_ik(const double* pdes, const double* qini, int leg, double* q, double* J) {
  Kinematicclass k; put3(k.Inverse_kinematics(get3(pdes), get3(qini), leg), q); putJ(k, J);
}
void ref_ik_g(const double* bp, const double* br, const double* pdes, const double* qini, int leg,
              double* q, double* J) {
  Kinematicclass k; put3(k.Inverse_kinematics_g(get3(bp), get3(br), get3(pdes), get3(qini), leg), q); putJ(k, J);
}

}  // extern "C"
