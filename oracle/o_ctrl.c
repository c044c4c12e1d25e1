/* TEST INFRASTRUCTURE - CPU oracle, controller half of the path (pinned against the reference's own Python, see
 * tools/gen_osc_golden.py -> tests/golden/osc_golden.npz):
 *   OperationalSpaceController.set_goal / run_controller  robosuite/controllers/parts/arm/osc.py:225-283,403-495
 *   opspace_matrices / nullspace_torques / orientation_error  robosuite/utils/control_utils.py:7-111
 *   Controller.scale_action                                robosuite/controllers/parts/controller.py:149-168
 *   PandaGripper.format_action                             robosuite/models/grippers/panda_gripper.py:43-58
 *   SimpleGripController.run_controller                    robosuite/controllers/parts/gripper/simple_grip.py:150-186
 *   FixedBaseRobot.control (clip to ctrlrange, write ctrl) robosuite/robots/fixed_base_robot.py:121-153
 *   MujocoEnv.step substep loop                            robosuite/environments/base.py:494-505 */
#include "b2s_oracle.h"
#include "o_math.h"
#include <stdlib.h>

typedef struct {
  int kind, action_dim, n_arm;
  int arm_dof[8], arm_qpos[8], arm_act[8];
  int eef_site, base_site, n_grip;
  int grip_act[4];
  double grip_sign[4], grip_speed;
  double kp[6], damping_ratio[6], input_max[6], input_min[6], output_max[6], output_min[6];
  double null_kp;
  int uncouple_pos_ori, n_obs_site;
  /* JOINT_VELOCITY (controllers/parts/generic/joint_vel.py:60-209): per-joint PID gains and action scaling */
  double jv_kp[8], jv_ki[8], jv_kd[8], jv_in_max[8], jv_in_min[8], jv_out_max[8], jv_out_min[8];
  double jv_vel_lo, jv_vel_hi;
  int jv_use_vel_limits, jv_torque_comp;
} OCtrlCfg; /* same layout as b2s_ctrl_cfg in include/b2s.h */

typedef struct {
  double goal_pos[3], goal_ori[9], initial_joint[8], grip_action[4];
  double torques[8]; /* last arm torques before clipping */
  /* JOINT_VELOCITY state (joint_vel.py:104-110) */
  double jv_goal[8], jv_last_err[8], jv_summed[8], jv_derr[5][8];
  int jv_ptr, jv_size, jv_saturated;
} OCtrlState;

/* small dense helpers (n <= 8) */
static int inv_spd(double* A, int n) { /* in-place inverse via Gauss-Jordan with partial pivoting */
  double B[64];
  for (int i = 0; i < n * n; i++) B[i] = 0;
  for (int i = 0; i < n; i++) B[i * n + i] = 1;
  for (int c = 0; c < n; c++) {
    int piv = c;
    for (int r = c + 1; r < n; r++) if (fabs(A[r * n + c]) > fabs(A[piv * n + c])) piv = r;
    if (fabs(A[piv * n + c]) < 1e-300) return -1;
    if (piv != c)
      for (int k = 0; k < n; k++) {
        double t = A[c * n + k]; A[c * n + k] = A[piv * n + k]; A[piv * n + k] = t;
        t = B[c * n + k]; B[c * n + k] = B[piv * n + k]; B[piv * n + k] = t;
      }
    double inv = 1.0 / A[c * n + c];
    for (int k = 0; k < n; k++) { A[c * n + k] *= inv; B[c * n + k] *= inv; }
    for (int r = 0; r < n; r++) {
      if (r == c) continue;
      double f = A[r * n + c];
      if (f == 0) continue;
      for (int k = 0; k < n; k++) { A[r * n + k] -= f * A[c * n + k]; B[r * n + k] -= f * B[c * n + k]; }
    }
  }
  memcpy(A, B, sizeof(double) * n * n);
  return 0;
}

/* Moore-Penrose inverse of a symmetric PSD matrix via Jacobi eigen-decomposition with numpy.linalg.pinv's default
 * cutoff rcond = 1e-15 * max singular value (control_utils.py:74-76) */
static void pinv_sym(double* A, int n) {
  double V[36], D[36];
  memcpy(D, A, sizeof(double) * n * n);
  for (int i = 0; i < n * n; i++) V[i] = 0;
  for (int i = 0; i < n; i++) V[i * n + i] = 1;
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0;
    for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) off += D[p * n + q] * D[p * n + q];
    if (off < 1e-300) break;
    for (int p = 0; p < n; p++)
      for (int q = p + 1; q < n; q++) {
        if (fabs(D[p * n + q]) < 1e-300) continue;
        double theta = (D[q * n + q] - D[p * n + p]) / (2 * D[p * n + q]);
        double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
        double c = 1 / sqrt(t * t + 1), s = t * c;
        for (int k = 0; k < n; k++) {
          double a = D[k * n + p], b = D[k * n + q];
          D[k * n + p] = c * a - s * b; D[k * n + q] = s * a + c * b;
        }
        for (int k = 0; k < n; k++) {
          double a = D[p * n + k], b = D[q * n + k];
          D[p * n + k] = c * a - s * b; D[q * n + k] = s * a + c * b;
        }
        for (int k = 0; k < n; k++) {
          double a = V[k * n + p], b = V[k * n + q];
          V[k * n + p] = c * a - s * b; V[k * n + q] = s * a + c * b;
        }
      }
  }
  double smax = 0;
  for (int i = 0; i < n; i++) if (fabs(D[i * n + i]) > smax) smax = fabs(D[i * n + i]);
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) {
      double s = 0;
      for (int k = 0; k < n; k++) {
        double ev = D[k * n + k];
        if (fabs(ev) > 1e-15 * smax) s += V[i * n + k] * V[j * n + k] / ev;
      }
      A[i * n + j] = s;
    }
}

/* delta rotation with the reference's float32 round trip (transform_utils.py:461-487, 515-538) */
static void delta_rotmat(double* Rm, const double* aa) {
  double angle = sqrt(aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2]);
  double qd[4] = {0, 0, 0, 1}; /* x y z w */
  if (angle != 0.0) { /* math.isclose(angle, 0.0) uses rel_tol only: true for exactly 0 */
    double s = sin(angle / 2.0);
    qd[0] = aa[0] / angle * s; qd[1] = aa[1] / angle * s; qd[2] = aa[2] / angle * s; qd[3] = cos(angle / 2.0);
  }
  float q[4] = {(float)qd[3], (float)qd[0], (float)qd[1], (float)qd[2]}; /* w x y z, cast to float32 */
  float n = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  if (n < 8.881784197001252e-16f) { /* EPS = finfo(float).eps * 4 */
    for (int i = 0; i < 9; i++) Rm[i] = (i % 4 == 0);
    return;
  }
  float rr = 2.0f / n;
  float sc = (float)sqrt((double)rr);
  for (int i = 0; i < 4; i++) q[i] *= sc;
  float q2[4][4];
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) q2[i][j] = q[i] * q[j];
  Rm[0] = 1.0f - q2[2][2] - q2[3][3]; Rm[1] = q2[1][2] - q2[3][0]; Rm[2] = q2[1][3] + q2[2][0];
  Rm[3] = q2[1][2] + q2[3][0]; Rm[4] = 1.0f - q2[1][1] - q2[3][3]; Rm[5] = q2[2][3] - q2[1][0];
  Rm[6] = q2[1][3] - q2[2][0]; Rm[7] = q2[2][3] + q2[1][0]; Rm[8] = 1.0f - q2[1][1] - q2[2][2];
}

/* controller.reset_goal + initial joints (osc.py:520-544, controller.py:126-132); requires a prior o_forward */
void o_ctrl_reset(const OModel* m, const OData* d, const OCtrlCfg* c, OCtrlState* s) {
  memcpy(s->goal_pos, d->site_xpos + 3 * c->eef_site, 3 * sizeof(double));
  memcpy(s->goal_ori, d->site_xmat + 9 * c->eef_site, 9 * sizeof(double));
  for (int i = 0; i < c->n_arm; i++) s->initial_joint[i] = d->qpos[c->arm_qpos[i]];
  for (int i = 0; i < 4; i++) s->grip_action[i] = 0;
  for (int i = 0; i < 8; i++) s->torques[i] = 0;
  memset(s->jv_goal, 0, sizeof s->jv_goal); memset(s->jv_last_err, 0, sizeof s->jv_last_err);
  memset(s->jv_summed, 0, sizeof s->jv_summed); memset(s->jv_derr, 0, sizeof s->jv_derr);
  s->jv_ptr = 4; s->jv_size = 0; s->jv_saturated = 0;
  if (c->kind == 3) /* JointPositionController.reset_goal: goal <- current joint positions */
    for (int i = 0; i < c->n_arm; i++) s->jv_goal[i] = d->qpos[c->arm_qpos[i]];
}

static void grip_run(const OModel* m, OData* d, const OCtrlCfg* c, OCtrlState* s, const double* action, int grip_index) {
  if (action)
    for (int g = 0; g < c->n_grip; g++) {
      double a = action[grip_index];
      double sg = a > 0 ? 1.0 : (a < 0 ? -1.0 : 0.0);
      s->grip_action[g] = fmin(fmax(s->grip_action[g] + c->grip_sign[g] * c->grip_speed * sg, -1.0), 1.0);
    }
  for (int g = 0; g < c->n_grip; g++) {
    int u = c->grip_act[g];
    double lo = m->actuator_ctrlrange[2 * u], hi = m->actuator_ctrlrange[2 * u + 1];
    double v = 0.5 * (hi + lo) + 0.5 * (hi - lo) * s->grip_action[g];
    d->ctrl[u] = fmin(fmax(v, lo), hi);
  }
}

/* JointVelocityController.set_goal / run_controller (joint_vel.py:129-209), with the constructor's broken
 * `self.torque_compensation = ...` (:127, assigns to a read-only property) read as the sibling controllers spell it
 * (joint_tor.py:109 `use_torque_compensation`): PID on joint velocity + qfrc_bias, clipped, anti-windup on saturation */
static void jv_run(const OModel* m, OData* d, const OCtrlCfg* c, OCtrlState* s, const double* action) {
  int na = c->n_arm;
  if (action)
    for (int k = 0; k < na; k++) {
      double a = fmin(fmax(action[k], c->jv_in_min[k]), c->jv_in_max[k]);
      double scale = fabs(c->jv_out_max[k] - c->jv_out_min[k]) / fabs(c->jv_in_max[k] - c->jv_in_min[k]);
      double g = (a - 0.5 * (c->jv_in_max[k] + c->jv_in_min[k])) * scale + 0.5 * (c->jv_out_max[k] + c->jv_out_min[k]);
      if (c->jv_use_vel_limits) g = fmin(fmax(g, c->jv_vel_lo), c->jv_vel_hi);
      s->jv_goal[k] = g;
    }
  s->jv_ptr = (s->jv_ptr + 1) % 5;
  if (s->jv_size < 5) s->jv_size++;
  double diff = 0;
  for (int k = 0; k < na; k++) {
    double err = s->jv_goal[k] - d->qvel[c->arm_dof[k]];
    s->jv_derr[s->jv_ptr][k] = err - s->jv_last_err[k];
    s->jv_last_err[k] = err;
    if (!s->jv_saturated) s->jv_summed[k] += err;
    double avg = 0;
    for (int r = 0; r < s->jv_size; r++) avg += s->jv_derr[r][k];
    avg /= s->jv_size;
    double tau = c->jv_kp[k] * err + c->jv_ki[k] * s->jv_summed[k] + c->jv_kd[k] * avg;
    if (c->jv_torque_comp) tau += d->qfrc_bias[c->arm_dof[k]];
    int u = c->arm_act[k];
    double cl = fmin(fmax(tau, m->actuator_ctrlrange[2 * u]), m->actuator_ctrlrange[2 * u + 1]);
    s->torques[k] = tau;
    d->ctrl[u] = cl;
    diff += fabs(cl - tau);
  }
  s->jv_saturated = diff != 0;
  grip_run(m, d, c, s, action, na);
}

/* scale_action of the joint-space controllers (controller.py:149-168) with the per-joint limits kept in the jv_* fields */
static double joint_scale(const OCtrlCfg* c, int k, double a) {
  a = fmin(fmax(a, c->jv_in_min[k]), c->jv_in_max[k]);
  double scale = fabs(c->jv_out_max[k] - c->jv_out_min[k]) / fabs(c->jv_in_max[k] - c->jv_in_min[k]);
  return (a - 0.5 * (c->jv_in_max[k] + c->jv_in_min[k])) * scale + 0.5 * (c->jv_out_max[k] + c->jv_out_min[k]);
}

/* JointPositionController (joint_pos.py:160-262), input_type "delta", impedance_mode "fixed", no interpolator:
 * goal_qpos = joint_pos + scaled delta at policy steps; torque = M_arm (kp e - kd qvel) + qfrc_bias, clipped.
 * kp / kd live in jv_kp / jv_kd, the goal in jv_goal. */
static void jp_run(const OModel* m, OData* d, const OCtrlCfg* c, OCtrlState* s, const double* action) {
  int na = c->n_arm, nv = m->nv;
  if (action)
    for (int k = 0; k < na; k++) s->jv_goal[k] = d->qpos[c->arm_qpos[k]] + joint_scale(c, k, action[k]);
  double des[8];
  for (int k = 0; k < na; k++)
    des[k] = (s->jv_goal[k] - d->qpos[c->arm_qpos[k]]) * c->jv_kp[k] - d->qvel[c->arm_dof[k]] * c->jv_kd[k];
  for (int a = 0; a < na; a++) {
    double tau = 0;
    if (c->jv_torque_comp) {
      for (int b = 0; b < na; b++) tau += d->M[c->arm_dof[a] * nv + c->arm_dof[b]] * des[b];
      tau += d->qfrc_bias[c->arm_dof[a]];
    } else tau = des[a];
    int u = c->arm_act[a];
    s->torques[a] = tau;
    d->ctrl[u] = fmin(fmax(tau, m->actuator_ctrlrange[2 * u]), m->actuator_ctrlrange[2 * u + 1]);
  }
  grip_run(m, d, c, s, action, na);
}

/* JointTorqueController (joint_tor.py:112-160): goal_torque = clip(scaled action, actuator limits) at policy steps;
 * torque = goal_torque + qfrc_bias, clipped.  The goal lives in jv_goal. */
static void jt_run(const OModel* m, OData* d, const OCtrlCfg* c, OCtrlState* s, const double* action) {
  int na = c->n_arm;
  if (action)
    for (int k = 0; k < na; k++) {
      int u = c->arm_act[k];
      s->jv_goal[k] = fmin(fmax(joint_scale(c, k, action[k]), m->actuator_ctrlrange[2 * u]), m->actuator_ctrlrange[2 * u + 1]);
    }
  for (int k = 0; k < na; k++) {
    int u = c->arm_act[k];
    double tau = s->jv_goal[k] + (c->jv_torque_comp ? d->qfrc_bias[c->arm_dof[k]] : 0.0);
    s->torques[k] = tau;
    d->ctrl[u] = fmin(fmax(tau, m->actuator_ctrlrange[2 * u]), m->actuator_ctrlrange[2 * u + 1]);
  }
  grip_run(m, d, c, s, action, na);
}

/* one controller evaluation between step1 and step2; action != NULL on policy steps */
void o_ctrl_run(const OModel* m, OData* d, const OCtrlCfg* c, OCtrlState* s, const double* action) {
  if (c->kind == 2) { jv_run(m, d, c, s, action); return; }
  if (c->kind == 3) { jp_run(m, d, c, s, action); return; }
  if (c->kind == 4) { jt_run(m, d, c, s, action); return; }
  int nv = m->nv, na = c->n_arm;
  const double* ref_pos = d->site_xpos + 3 * c->eef_site;
  const double* ref_ori = d->site_xmat + 9 * c->eef_site;
  const double* org_pos = d->site_xpos + 3 * c->base_site;
  const double* org_ori = d->site_xmat + 9 * c->base_site;
  if (action) {
    /* scale_action: clip to [input_min, input_max], affine map to [output_min, output_max] */
    /* OSC_POSITION (kind 5, osc.py:152-166, 259-270): 3-dim arm action, the orientation part of the delta is zero, so every
     * policy step re-anchors goal_ori to the current orientation */
    double sd[6] = {0, 0, 0, 0, 0, 0};
    int od = c->kind == 5 ? 3 : 6;
    for (int k = 0; k < od; k++) {
      double a = fmin(fmax(action[k], c->input_min[k]), c->input_max[k]);
      double scale = fabs(c->output_max[k] - c->output_min[k]) / fabs(c->input_max[k] - c->input_min[k]);
      sd[k] = (a - 0.5 * (c->input_max[k] + c->input_min[k])) * scale + 0.5 * (c->output_max[k] + c->output_min[k]);
    }
    /* goal_pos = R_o^T (ref_pos - o) + delta ; goal_ori = R(delta) R_o^T R_ref  (achieved mode, base frame) */
    double rel[3], inb[3], cur[9], Rd[9];
    v3_sub(rel, ref_pos, org_pos);
    m3_mulTv(inb, org_ori, rel);
    for (int k = 0; k < 3; k++) s->goal_pos[k] = inb[k] + sd[k];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) cur[3 * i + j] = org_ori[i] * ref_ori[j] + org_ori[3 + i] * ref_ori[3 + j] + org_ori[6 + i] * ref_ori[6 + j];
    delta_rotmat(Rd, sd + 3);
    m3_mul(s->goal_ori, Rd, cur);
    /* gripper: format_action integrates sign(a)*speed into current_action, clipped to [-1,1] */
    for (int g = 0; g < c->n_grip; g++) {
      double a = action[od];
      double sg = a > 0 ? 1.0 : (a < 0 ? -1.0 : 0.0);
      s->grip_action[g] = fmin(fmax(s->grip_action[g] + c->grip_sign[g] * c->grip_speed * sg, -1.0), 1.0);
    }
  }
  /* site Jacobians restricted to the arm dofs, site velocities */
  double* jp = (double*)malloc(sizeof(double) * 6 * nv);
  double* jr = jp + 3 * nv;
  o_jac(m, d, jp, jr, ref_pos, m->site_bodyid[c->eef_site]);
  double J[6 * 8], vel[6] = {0, 0, 0, 0, 0, 0}, bvel[6] = {0, 0, 0, 0, 0, 0};
  for (int r = 0; r < 3; r++)
    for (int i = 0; i < nv; i++) { vel[r] += jp[r * nv + i] * d->qvel[i]; vel[3 + r] += jr[r * nv + i] * d->qvel[i]; }
  for (int r = 0; r < 3; r++)
    for (int k = 0; k < na; k++) { J[r * na + k] = jp[r * nv + c->arm_dof[k]]; J[(3 + r) * na + k] = jr[r * nv + c->arm_dof[k]]; }
  o_jac(m, d, jp, jr, org_pos, m->site_bodyid[c->base_site]);
  for (int r = 0; r < 3; r++)
    for (int i = 0; i < nv; i++) { bvel[r] += jp[r * nv + i] * d->qvel[i]; bvel[3 + r] += jr[r * nv + i] * d->qvel[i]; }
  free(jp);
  /* desired pose in world, errors, desired wrench */
  double des_pos[3], des_ori[9], err[6], F[6];
  m3_mulv(des_pos, org_ori, s->goal_pos);
  v3_add(des_pos, des_pos, org_pos);
  m3_mul(des_ori, org_ori, s->goal_ori);
  v3_sub(err, des_pos, ref_pos);
  double e3[3] = {0, 0, 0};
  for (int col = 0; col < 3; col++) {
    double rc[3] = {ref_ori[col], ref_ori[3 + col], ref_ori[6 + col]}, rd[3] = {des_ori[col], des_ori[3 + col], des_ori[6 + col]}, cr[3];
    v3_cross(cr, rc, rd);
    v3_add(e3, e3, cr);
  }
  for (int k = 0; k < 3; k++) err[3 + k] = 0.5 * e3[k];
  for (int k = 0; k < 6; k++) {
    double kd = 2 * sqrt(c->kp[k]) * c->damping_ratio[k];
    F[k] = err[k] * c->kp[k] + (-(vel[k] - bvel[k])) * kd;
  }
  /* operational-space matrices */
  double Mm[64], Mi[64], MiJt[8 * 6], Lf[36], Lp[9], Lo[9];
  for (int a = 0; a < na; a++)
    for (int b = 0; b < na; b++) Mm[a * na + b] = d->M[c->arm_dof[a] * nv + c->arm_dof[b]];
  memcpy(Mi, Mm, sizeof(double) * na * na);
  inv_spd(Mi, na);
  for (int a = 0; a < na; a++)
    for (int r = 0; r < 6; r++) {
      double sacc = 0;
      for (int b = 0; b < na; b++) sacc += Mi[a * na + b] * J[r * na + b];
      MiJt[a * 6 + r] = sacc;
    }
  for (int r = 0; r < 6; r++)
    for (int q = 0; q < 6; q++) {
      double sacc = 0;
      for (int a = 0; a < na; a++) sacc += J[r * na + a] * MiJt[a * 6 + q];
      Lf[r * 6 + q] = sacc;
    }
  for (int r = 0; r < 3; r++)
    for (int q = 0; q < 3; q++) { Lp[r * 3 + q] = Lf[r * 6 + q]; Lo[r * 3 + q] = Lf[(3 + r) * 6 + 3 + q]; }
  pinv_sym(Lf, 6);
  pinv_sym(Lp, 3);
  pinv_sym(Lo, 3);
  double W[6];
  if (c->uncouple_pos_ori) {
    for (int r = 0; r < 3; r++) {
      W[r] = Lp[r * 3] * F[0] + Lp[r * 3 + 1] * F[1] + Lp[r * 3 + 2] * F[2];
      W[3 + r] = Lo[r * 3] * F[3] + Lo[r * 3 + 1] * F[4] + Lo[r * 3 + 2] * F[5];
    }
  } else {
    for (int r = 0; r < 6; r++) { W[r] = 0; for (int q = 0; q < 6; q++) W[r] += Lf[r * 6 + q] * F[q]; }
  }
  /* nullspace: N = I - Jbar J, Jbar = M^-1 J^T Lambda_full ; tau_null = N^T M (kp (q0 - q) - kv qdot) */
  double Jbar[8 * 6], N[64], pt[8], ptm[8];
  for (int a = 0; a < na; a++)
    for (int q = 0; q < 6; q++) { double sacc = 0; for (int r = 0; r < 6; r++) sacc += MiJt[a * 6 + r] * Lf[r * 6 + q]; Jbar[a * 6 + q] = sacc; }
  for (int a = 0; a < na; a++)
    for (int b = 0; b < na; b++) { double sacc = 0; for (int r = 0; r < 6; r++) sacc += Jbar[a * 6 + r] * J[r * na + b]; N[a * na + b] = (a == b) - sacc; }
  double kv = 2 * sqrt(c->null_kp);
  for (int a = 0; a < na; a++) pt[a] = c->null_kp * (s->initial_joint[a] - d->qpos[c->arm_qpos[a]]) - kv * d->qvel[c->arm_dof[a]];
  for (int a = 0; a < na; a++) { double sacc = 0; for (int b = 0; b < na; b++) sacc += Mm[a * na + b] * pt[b]; ptm[a] = sacc; }
  for (int a = 0; a < na; a++) {
    double tau = d->qfrc_bias[c->arm_dof[a]];
    for (int r = 0; r < 6; r++) tau += J[r * na + a] * W[r];
    for (int b = 0; b < na; b++) tau += N[b * na + a] * ptm[b];
    s->torques[a] = tau;
    int u = c->arm_act[a];
    d->ctrl[u] = fmin(fmax(tau, m->actuator_ctrlrange[2 * u]), m->actuator_ctrlrange[2 * u + 1]);
  }
  for (int g = 0; g < c->n_grip; g++) {
    int u = c->grip_act[g];
    double lo = m->actuator_ctrlrange[2 * u], hi = m->actuator_ctrlrange[2 * u + 1];
    double v = 0.5 * (hi + lo) + 0.5 * (hi - lo) * s->grip_action[g];
    d->ctrl[u] = fmin(fmax(v, lo), hi);
  }
}

/* MujocoEnv.step substep loop (environments/base.py:494-505), lite_physics=True */
void o_env_step(const OModel* m, OData* d, const OCtrlCfg* c, OCtrlState* s, const double* action, int nsub) {
  for (int i = 0; i < nsub; i++) {
    o_step1(m, d);
    o_ctrl_run(m, d, c, s, i == 0 ? action : NULL);
    o_step2(m, d);
  }
}
