// Host build of the controller arithmetic the device runs (robosuite_b200/csrc/b2s_oscmath.h), for tests/test_oscmath_host.py.
#include "../../robosuite_b200/csrc/b2s_oscmath.h"

extern "C" void osc_host_torques(int na, const double* cdof_arm, const double* ref_pos, const double* ref_ori, const double* org_pos,
                                 const double* org_ori, const double* goal_pos, const double* goal_ori, const double* cvel_eef,
                                 const double* cvel_base, const double* Marm, const double* bias, const double* qpos_arm,
                                 const double* qvel_arm, const double* init_q, const double* kp, const double* kd, double null_kp,
                                 int uncouple, double* tau) {
  float Jf[6 * OSC_NA_MAX];
  double work[OSC_WORK_DOUBLES];
  for (int a = 0; a < na; a++) {
    double col[6];
    osc_jac_col(cdof_arm + 6 * a, ref_pos, col);
    for (int r = 0; r < 6; r++) Jf[r * na + a] = (float)col[r];
  }
  // keep full double precision in the host check: use a double J instead of the float tile
  double Jd[6 * OSC_NA_MAX];
  for (int a = 0; a < na; a++) {
    double col[6];
    osc_jac_col(cdof_arm + 6 * a, ref_pos, col);
    for (int r = 0; r < 6; r++) Jd[r * na + a] = col[r];
  }
  for (int a = 0; a < na; a++)
    for (int b = 0; b <= a; b++) work[OSC_OFF_L + osc_tri(a, b)] = Marm[a * na + b];
  double vel[6], bvel[6], F[6], pt[OSC_NA_MAX];
  osc_jac_col(cvel_eef, ref_pos, vel);
  osc_jac_col(cvel_base, org_pos, bvel);
  osc_wrench(ref_pos, ref_ori, org_pos, org_ori, goal_pos, goal_ori, vel, bvel, kp, kd, F);
  double kv = 2.0 * sqrt(null_kp);
  for (int a = 0; a < na; a++) pt[a] = null_kp * (init_q[a] - qpos_arm[a]) - kv * qvel_arm[a];
  osc_torques(OscView<double>{Jd, 1}, OscView<double>{work, 1}, na, F, pt, bias, uncouple, tau);
  (void)Jf;
}
