// Exercises the C++ mirror of the reference classes (quadrupedal_loco_b200/host/go1mpc.hpp)
// the way the reference's callers use them, and prints the results as JSON for
// tests/test_host_cpp.py to compare with the CPU oracle.  Needs a GPU.
#include <cstdio>
#include <cmath>
#include "../../quadrupedal_loco_b200/host/go1mpc.hpp"

using namespace go1host;

static void arr(const char* name, const double* v, int n, bool last = false) {
  printf("\"%s\": [", name);
  for (int i = 0; i < n; i++) printf("%s%.17g", i ? ", " : "", v[i]);
  printf("]%s\n", last ? "" : ",");
}

int main() {
  printf("{\n");
  // --- QPBaseClass seam: the QuadProg++ demo problem
  QPBase qp;
  qp.resizeQP(2, 1, 3);
  qp.G(0, 0) = 4; qp.G(0, 1) = -2; qp.G(1, 0) = -2; qp.G(1, 1) = 4;
  qp._g0[0] = 6; qp._g0[1] = 0;
  qp.CE(0, 0) = 1; qp.CE(1, 0) = 1; qp._ce0[0] = -3;
  qp.CI(0, 0) = 1; qp.CI(1, 0) = 0; qp.CI(0, 1) = 0; qp.CI(1, 1) = 1; qp.CI(0, 2) = 1; qp.CI(1, 2) = 1;
  qp._ci0[0] = 0; qp._ci0[1] = 0; qp._ci0[2] = -2;
  bool ok = qp.solveQP();
  printf("\"qp_ok\": %d, \"qp_cost\": %.17g,\n", ok ? 1 : 0, qp._cost);
  arr("qp_x", qp._X.data(), 2);

  // --- PRMPCClass::body_theta_mpc, horizon 10, 30 closed-loop ticks
  const int nh = 10;
  BodyInclinationMPC body(nh);
  body._thetaxk(0) = 0.05; body._thetaxk(1) = -0.4; body._thetayk(0) = -0.08; body._thetayk(1) = 0.6;
  MatX zmp(2, nh), ang(2, nh), rf(2, nh), lf(2, nh), ca(3, nh);
  for (int k = 0; k < nh; k++) {
    zmp(0, k) = 0.30 + 0.001 * k; zmp(1, k) = 0.12; ang(0, k) = 0.21 - 0.002 * k; ang(1, k) = -0.19;
    rf(0, k) = 0.28; rf(1, k) = -0.127; lf(0, k) = 0.31; lf(1, k) = 0.126; ca(2, k) = 0.3 * std::sin(0.7 * k);
  }
  Vec<4> meas; Vec<9> nrt;
  double o14[30 * 14], th[30 * 4];
  int idx[30 * 2];
  for (int t = 0; t < 30; t++) {
    Vec<14> o = body.body_theta_mpc(200 + t, meas, zmp, ang, rf, lf, ca, nrt);
    for (int k = 0; k < 14; k++) o14[t * 14 + k] = o(k);
    th[t * 4] = body._thetaxk(0); th[t * 4 + 1] = body._thetaxk(1); th[t * 4 + 2] = body._thetayk(0); th[t * 4 + 3] = body._thetayk(1);
    idx[t * 2] = body._bjx1; idx[t * 2 + 1] = body._bjx2;
  }
  arr("body_out14", o14, 30 * 14); arr("body_theta", th, 30 * 4);
  printf("\"body_idx\": ["); for (int i = 0; i < 60; i++) printf("%s%d", i ? ", " : "", idx[i]); printf("],\n");

  // --- NLPClass::step_timing_opti_loop, the first 120 ticks of the replay
  StepTimingMPC nlp;
  nlp.FootStepInputs(0.2535, 0.075, 0.0); nlp.Initialize();
  Vec<18> est; Vec<3> rfb, lfb; rfb(1) = -0.12675; lfb(1) = 0.12675;
  double o38[120 * 38], o18[120 * 18];
  int rs[120];
  for (int i = 1; i <= 120; i++) {
    Vec<38> o = nlp.step_timing_opti_loop(i, est, rfb, lfb, 0.0, false);
    for (int k = 0; k < 38; k++) o38[(i - 1) * 38 + k] = o(k);
    Vec<18> f = nlp.Foot_trajectory_solve_mod2(i, false);
    for (int k = 0; k < 18; k++) o18[(i - 1) * 18 + k] = f(k);
    rs[i - 1] = nlp.right_support;
  }
  arr("step_out38", o38, 120 * 38); arr("foot_out18", o18, 120 * 18);
  printf("\"right_support\": ["); for (int i = 0; i < 120; i++) printf("%s%d", i ? ", " : "", rs[i]); printf("],\n");

  // --- Dynamiccclass: force_distribution -> force_opt -> compute_joint_torques as go1_servo chains them (servo.cpp:1200-1243)
  {
    GrfDistributor dyn;
    Vec<3> com; com(0) = 0.01; com(1) = -0.02; com(2) = 0.31;
    const double hom[12] = {0.1881, -0.1268, 0, 0.1881, 0.1268, 0, -0.1881, -0.1268, 0, -0.1881, 0.1268, 0};
    Vec<12> legs; for (int k = 0; k < 12; k++) legs(k) = hom[k] + 0.01 * std::sin(1.3 * k);
    Vec<6> F6; F6(0) = 3; F6(1) = -2; F6(2) = 60; F6(3) = -4; F6(4) = 5; F6(5) = 55;
    Vec<6> FT; FT(0) = 5; FT(1) = -7; FT(2) = 117.6; FT(3) = 1; FT(4) = -2; FT(5) = 0.5;
    double rf[3] = {0.01, -0.127, 0}, lf[3] = {-0.01, 0.127, 0};
    double Fg[2 * 12], grf[2 * 12], tau[2 * 12];
    int okk[2];
    for (int c = 0; c < 2; c++) {
      const int mode = 101 + c, rs = c;       // pace / trot; second call starts from the first call's grf_opt
      dyn.force_distribution(com, legs, F6, mode, 0.9, rf, lf);
      Vec<3> p[4]; for (int l = 0; l < 4; l++) for (int k = 0; k < 3; k++) p[l](k) = legs(3 * l + k);
      dyn.force_opt(com, p[0], p[1], p[2], p[3], FT, mode, rs, 0.9);
      for (int k = 0; k < 12; k++) { Fg[12 * c + k] = dyn.F_leg_guess(k); grf[12 * c + k] = dyn.grf_opt(k); }
      okk[c] = dyn.qp_solution ? 1 : 0;
      for (int l = 0; l < 4; l++) {
        Mat<3, 3> J;
        for (int r = 0; r < 3; r++) for (int k = 0; k < 3; k++) J(r, k) = 0.1 * std::cos(0.9 * (3 * r + k) + l);
        Vec<3> pd, pe, vd, ve;
        for (int k = 0; k < 3; k++) { pd(k) = 0.1 * k; pe(k) = 0.1 * k + 0.01; vd(k) = 0.2; ve(k) = 0.15 - 0.1 * k; }
        Vec<3> t = dyn.compute_joint_torques(J, (l + c) % 2 == 0, pd, pe, vd, ve, l);
        for (int k = 0; k < 3; k++) tau[12 * c + 3 * l + k] = t(k);
      }
    }
    arr("grf_guess", Fg, 24); arr("grf_opt", grf, 24); arr("grf_tau", tau, 24);
    printf("\"grf_ok\": [%d, %d],\n", okk[0], okk[1]);
  }

  // --- the two MPC nodes in lock step, as mpc.cpp / gait_fast.cpp run them: NLPRTControlClass::WalkingReactStepping every 25 ms
  //     (with a StopWalking at slow tick 260), the 100 Hz node consuming the latest message every 10 ms; first 400 slow ticks
  {
    GaitPlannerNode slow;
    FastGaitNode fast(4);
    Vec<18> est2; Vec<3> rf2, lf2; Vec<4> bs;
    Vec<100> msg;
    const int n_slow = 400;
    static double slow_msgs[400 * 100], fast_msgs[1001 * 100];
    int count = 0, nfast = 0;
    for (int t_ms = 0; count < n_slow || t_ms % 25; t_ms += 5) {
      if (t_ms % 25 == 0) {
        count++;
        if (count == 260) slow.StopWalking();
        msg = slow.WalkingReactStepping(count, true, est2, rf2, lf2);
        for (int k = 0; k < 100; k++) slow_msgs[(count - 1) * 100 + k] = msg(k);
      }
      if (t_ms % 10 == 0 && nfast < 1001) {
        Vec<100> o = fast.tick(msg, true, bs);
        for (int k = 0; k < 100; k++) fast_msgs[nfast * 100 + k] = o(k);
        nfast++;
      }
    }
    arr("node_slow", slow_msgs, n_slow * 100); arr("node_fast", fast_msgs, nfast * 100);
    printf("\"node_nfast\": %d, \"node_right_support\": %d,\n", nfast, slow.right_support);
  }

  // --- Kinematicclass: FK_g -> IK_g round trip, Jacobian side channel
  LegKinematics kin;
  Vec<3> bp, br, q, qi; bp(2) = 0.31; br(0) = 0.05; br(1) = -0.04; br(2) = 0.1; q(0) = 0.1; q(1) = 0.8; q(2) = -1.5;
  qi(0) = 0.0; qi(1) = 0.87; qi(2) = -1.5;
  Vec<3> p = kin.Forward_kinematics_g(bp, br, q, 1);
  arr("fk_pos", p.v, 3); arr("fk_J", kin.Jacobian_kin.m, 9);
  Vec<3> qs = kin.Inverse_kinematics_g(bp, br, p, qi, 1);
  arr("ik_q", qs.v, 3);
  printf("\"ik_updates\": %d\n}\n", kin.ik_updates);
  return 0;
}
