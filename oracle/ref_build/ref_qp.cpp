// C-callable wrapper around the UNMODIFIED reference QP solver, compiled from
// /root/reference against oracle/eigen_shim.  Test infrastructure only: used to
// pin oracle/qp_gi.c (tests/test_oracle_vs_ref.py) and, optionally, as the CPU
// baseline ("kind": "reference").  `#define private public` exposes the
// solver's working-set members so the final active set can be read back; the
// reference sources themselves are not edited or copied.
#include <cmath>
#include <limits>
#define private public
#define protected public
#include <utils/EiQuadProg/EiQuadProg.hpp>
#include <QP/QPBaseClass.h>
#undef private
#undef protected

extern "C" {

// Eigen::QP::solve_quadprog, EiQuadProg.cpp:493-513.  Same outputs as
// orc_qp_solve; iteration counters are not observable in the reference.
int ref_qp_solve(int n, int p, int m, const double* G, const double* g0, const double* CE,
                 const double* ce0, const double* CI, const double* ci0, double* x, double* cost,
                 int* active, int* nactive) {
  Eigen::QP qp;
  qp.resize(n, p, m);
  for (int i = 0; i < m + p; i++) { qp._A(i) = 0; qp._iai(i) = 0; qp._u(i) = 0; }
  Eigen::MatrixXd Gm(n, n), CEm(n, p), CIm(n, m);
  Eigen::VectorXd g0v(n), ce0v(p), ci0v(m), xv(n);
  for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) Gm(i, j) = G[j * n + i];
  for (int j = 0; j < p; j++) for (int i = 0; i < n; i++) CEm(i, j) = CE[j * n + i];
  for (int j = 0; j < m; j++) for (int i = 0; i < n; i++) CIm(i, j) = CI[j * n + i];
  for (int i = 0; i < n; i++) { g0v(i) = g0[i]; xv(i) = x[i]; }
  for (int i = 0; i < p; i++) ce0v(i) = ce0[i];
  for (int i = 0; i < m; i++) ci0v(i) = ci0[i];
  double f = qp.solve_quadprog(Gm, g0v, CEm, ce0v, CIm, ci0v, xv);
  for (int i = 0; i < n; i++) x[i] = xv(i);
  *cost = f;
  // working set: equalities occupy A[0..p), active inequalities are exactly
  // those with iai == -1 (cpp:288-292, 464, 452-459); they sit at A[p..p+k).
  int k = 0;
  for (int i = 0; i < m; i++) if (qp._iai(i) == -1) k++;
  *nactive = p + k;
  for (int i = 0; i < p + k && i < m + p; i++) active[i] = qp._A(i);
  return 0;
}

// Through the reference's own container: QPBaseClass::resizeQP / solveQP,
// QPBaseClass.cpp:102-153.  Returns solveQP()'s bool.
int ref_qpbase_solve(int n, int p, int m, const double* G, const double* g0, const double* CE,
                     const double* ce0, const double* CI, const double* ci0, double* x) {
  QPBaseClass q;  // default backend name "EiQuadProg"
  q.resizeQP(n, p, m);
  for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) q._G(i, j) = G[j * n + i];
  for (int j = 0; j < p; j++) for (int i = 0; i < n; i++) q._CE(i, j) = CE[j * n + i];
  for (int j = 0; j < m; j++) for (int i = 0; i < n; i++) q._CI(i, j) = CI[j * n + i];
  for (int i = 0; i < n; i++) { q._g0(i) = g0[i]; q._X(i) = x[i]; }
  for (int i = 0; i < p; i++) q._ce0(i) = ce0[i];
  for (int i = 0; i < m; i++) q._ci0(i) = ci0[i];
  q._ineqRowIdx = 0;
  bool ok = q.solveQP();
  for (int i = 0; i < n; i++) x[i] = q._X(i);
  return ok ? 1 : 0;
}

}  // extern "C"
