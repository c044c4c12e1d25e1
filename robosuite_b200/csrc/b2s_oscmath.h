// Operational-space controller arithmetic for ONE environment, written so that the same source runs
//   * on the device with one THREAD per environment (ctrl_osc_kernel in b2s_ctrlkernel.cuh: every lane of a warp works on its own
//     environment, the per-thread work arrays are columns of a shared-memory tile, stride = 32), and
//   * on the host (tests/csrc/osc_host.cpp, stride = 1), where tests/test_oscmath_host.py checks it against the CPU oracle, which
//     is itself pinned to the reference's OperationalSpaceController (tests/golden/osc_golden.npz).
// Reference semantics, file:line -
//   OperationalSpaceController.run_controller        robosuite/controllers/parts/arm/osc.py:403-495
//   opspace_matrices (lambda = pinv(J M^-1 J^T), ...) robosuite/utils/control_utils.py:43-82
//   nullspace_torques                                  robosuite/utils/control_utils.py:7-40
//   orientation_error                                  robosuite/utils/control_utils.py:85-111
// All dense algebra is fp64 (Lambda^-1 = J M^-1 J^T is ill conditioned near arm singularities); the inputs are whatever the
// engine computed (fp32 in production).  `pinv` keeps numpy's meaning: singular values below 1e-15 * max are dropped; the fast
// path (Cholesky) is taken when the matrix is far from that cut-off, the exact path (Jacobi eigen-decomposition) otherwise.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define OSC_HD __host__ __device__ __forceinline__
#define OSC_NI __host__ __device__ __noinline__
#else
#define OSC_HD inline
#define OSC_NI
#endif

// strided element access: element i of a per-environment array lives at p[i * s]
template <typename T> struct OscView {
  T* p; int s;
  OSC_HD T& operator[](int i) const { return p[i * s]; }
  OSC_HD OscView<T> at(int off) const { return OscView<T>{p + off * s, s}; }
};
OSC_HD int osc_tri(int i, int j) { return i * (i + 1) / 2 + j; }  // packed lower triangle, i >= j

// Moore-Penrose inverse of a symmetric PSD n x n matrix (n <= 6, dense row-major A, thread-private) with numpy.linalg.pinv's
// default cut-off rcond = 1e-15 * largest singular value (control_utils.py:74-76): cyclic Jacobi eigen-decomposition.
OSC_NI void osc_pinv_sym_jacobi(double* A, int n) {
  double V[36], D[36];
  for (int i = 0; i < n * n; i++) { D[i] = A[i]; V[i] = 0; }
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

// in-place Cholesky of a packed symmetric matrix; the diagonal keeps 1 / l_jj.  Returns the smallest pivot divided by the largest
// diagonal entry of the input (<= 0 when a pivot was not positive).
template <typename V> OSC_HD double osc_chol_packed(V A, int n) {
  double dmax = 0, pmin = 1e300;
  for (int i = 0; i < n; i++) { double d = A[osc_tri(i, i)]; if (d > dmax) dmax = d; }
  for (int j = 0; j < n; j++) {
    double d = A[osc_tri(j, j)];
    for (int k = 0; k < j; k++) { double l = A[osc_tri(j, k)]; d -= l * l; }
    if (d < pmin) pmin = d;
    if (!(d > 1e-300)) d = 1e-300;
    double inv = 1.0 / sqrt(d);
    A[osc_tri(j, j)] = inv;
    for (int i = j + 1; i < n; i++) {
      double s = A[osc_tri(i, j)];
      for (int k = 0; k < j; k++) s -= A[osc_tri(i, k)] * A[osc_tri(j, k)];
      A[osc_tri(i, j)] = s * inv;
    }
  }
  return dmax > 0 ? pmin / dmax : -1.0;
}
// x <- (L L^T)^-1 x with the factor of osc_chol_packed
template <typename V, typename X> OSC_HD void osc_chol_solve(V L, int n, X x) {
  for (int i = 0; i < n; i++) {
    double s = x[i];
    for (int k = 0; k < i; k++) s -= L[osc_tri(i, k)] * x[k];
    x[i] = s * L[osc_tri(i, i)];
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = x[i];
    for (int k = i + 1; k < n; k++) s -= L[osc_tri(k, i)] * x[k];
    x[i] = s * L[osc_tri(i, i)];
  }
}

// w[0..2] <- pinv(A) f for the symmetric 3x3 block A = Lf[o..o+2][o..o+2] (packed 6x6 lower triangle Lf)
template <typename V> OSC_HD void osc_block3_apply(V Lf, int o, const double* f, double* w) {
  double a00 = Lf[osc_tri(o, o)], a01 = Lf[osc_tri(o + 1, o)], a02 = Lf[osc_tri(o + 2, o)];
  double a11 = Lf[osc_tri(o + 1, o + 1)], a12 = Lf[osc_tri(o + 2, o + 1)], a22 = Lf[osc_tri(o + 2, o + 2)];
  double c00 = a11 * a22 - a12 * a12, c01 = a02 * a12 - a01 * a22, c02 = a01 * a12 - a02 * a11;
  double c11 = a00 * a22 - a02 * a02, c12 = a01 * a02 - a00 * a12, c22 = a00 * a11 - a01 * a01;
  double det = a00 * c00 + a01 * c01 + a02 * c02;
  double tr = (a00 + a11 + a22) * (1.0 / 3.0);
  if (det > 1e-10 * tr * tr * tr) {  // well inside numpy's 1e-15 singular-value cut-off: plain inverse
    double id = 1.0 / det;
    w[0] = (c00 * f[0] + c01 * f[1] + c02 * f[2]) * id;
    w[1] = (c01 * f[0] + c11 * f[1] + c12 * f[2]) * id;
    w[2] = (c02 * f[0] + c12 * f[1] + c22 * f[2]) * id;
    return;
  }
  double A[9] = {a00, a01, a02, a01, a11, a12, a02, a12, a22};
  osc_pinv_sym_jacobi(A, 3);
  for (int r = 0; r < 3; r++) w[r] = A[3 * r] * f[0] + A[3 * r + 1] * f[1] + A[3 * r + 2] * f[2];
}

// ---- work-array layout (doubles per environment)
#define OSC_NA_MAX 8
#define OSC_OFF_L 0                      // packed na x na: arm mass matrix, then its Cholesky factor
#define OSC_OFF_LF 36                    // packed 6 x 6: lambda_full^-1 = J M^-1 J^T, then its Cholesky factor
#define OSC_OFF_X 57                     // na: solve vector
#define OSC_OFF_PTM 65                   // na: M (kp (q0 - q) - kv qdot)
#define OSC_OFF_Y 73                     // 6
#define OSC_WORK_DOUBLES 79

// Torques of one arm.  J: 6 x na (rows 0-2 translational, 3-5 rotational; element (r, a) at J[r * na + a]); work: OSC_WORK_DOUBLES
// doubles with work[OSC_OFF_L ...] holding the packed arm mass matrix on entry.  F: desired wrench before the lambda matrices,
// pt: nullspace posture input kp (q0 - q) - kv qdot, bias: qfrc_bias of the arm dofs.  tau: na outputs (before clipping).
template <typename JV, typename WV>
OSC_HD void osc_torques(JV J, WV work, int na, const double* F, const double* pt, const double* bias, int uncouple, double* tau) {
  WV L = work.at(OSC_OFF_L), Lf = work.at(OSC_OFF_LF), x = work.at(OSC_OFF_X), ptm = work.at(OSC_OFF_PTM), y = work.at(OSC_OFF_Y);
  // ptm = M pt (before M is factored in place)
  for (int a = 0; a < na; a++) {
    double s = 0;
    for (int b = 0; b < na; b++) s += L[a >= b ? osc_tri(a, b) : osc_tri(b, a)] * pt[b];
    ptm[a] = s;
  }
  osc_chol_packed(L, na);
  // lambda_full^-1 = J M^-1 J^T, one column at a time: x = M^-1 J_q^T
  for (int q = 0; q < 6; q++) {
    for (int a = 0; a < na; a++) x[a] = (double)J[q * na + a];
    osc_chol_solve(L, na, x);
    for (int r = q; r < 6; r++) {
      double s = 0;
      for (int a = 0; a < na; a++) s += (double)J[r * na + a] * x[a];
      Lf[osc_tri(r, q)] = s;
    }
  }
  // y = (M^-1 J^T)^T ptm = J M^-1 ptm
  for (int a = 0; a < na; a++) x[a] = ptm[a];
  osc_chol_solve(L, na, x);
  for (int r = 0; r < 6; r++) {
    double s = 0;
    for (int a = 0; a < na; a++) s += (double)J[r * na + a] * x[a];
    y[r] = s;
  }
  double W[6], z[6];
  if (uncouple) {  // decoupled wrench: lambda_pos F_pos, lambda_ori F_ori (osc.py:458-466)
    osc_block3_apply(Lf, 0, F, W);
    osc_block3_apply(Lf, 3, F + 3, W + 3);
  }
  // lambda_full = pinv(lambda_full^-1) applied to y (nullspace) and, coupled mode, to F.  Fast path: Cholesky in place; when a
  // pivot says the matrix may be within numpy's cut-off of singular, the exact path rebuilds it and eigen-decomposes.
  double ratio = osc_chol_packed(Lf, 6);
  if (ratio > 1e-11) {
    for (int r = 0; r < 6; r++) x[r] = y[r];
    osc_chol_solve(Lf, 6, x);
    for (int r = 0; r < 6; r++) z[r] = x[r];
    if (!uncouple) {
      for (int r = 0; r < 6; r++) x[r] = F[r];
      osc_chol_solve(Lf, 6, x);
      for (int r = 0; r < 6; r++) W[r] = x[r];
    }
  } else {
    double full[36];
    for (int q = 0; q < 6; q++) {
      for (int a = 0; a < na; a++) x[a] = (double)J[q * na + a];
      osc_chol_solve(L, na, x);
      for (int r = q; r < 6; r++) {
        double s = 0;
        for (int a = 0; a < na; a++) s += (double)J[r * na + a] * x[a];
        full[r * 6 + q] = s; full[q * 6 + r] = s;
      }
    }
    osc_pinv_sym_jacobi(full, 6);
    for (int r = 0; r < 6; r++) {
      double s = 0, t = 0;
      for (int q = 0; q < 6; q++) { s += full[r * 6 + q] * y[q]; t += full[r * 6 + q] * F[q]; }
      z[r] = s;
      if (!uncouple) W[r] = t;
    }
  }
  // tau = J^T W + N^T ptm + bias,  N^T ptm = ptm - J^T lambda_full (J M^-1 ptm)
  for (int a = 0; a < na; a++) {
    double s = bias[a] + ptm[a];
    for (int r = 0; r < 6; r++) s += (double)J[r * na + a] * (W[r] - z[r]);
    tau[a] = s;
  }
}

// column of the site Jacobian for one dof: cd = spatial motion axis [angular; linear at the world origin], p = site position.
// out[0..2] translational (v_O + w x p), out[3..5] rotational (binding_utils.py:826-852 get_site_jacp / jacr)
template <typename T> OSC_HD void osc_jac_col(const T* cd, const T* p, T* out) {
  out[0] = cd[3] + (cd[1] * p[2] - cd[2] * p[1]);
  out[1] = cd[4] + (cd[2] * p[0] - cd[0] * p[2]);
  out[2] = cd[5] + (cd[0] * p[1] - cd[1] * p[0]);
  out[3] = cd[0]; out[4] = cd[1]; out[5] = cd[2];
}

// Desired wrench F = kp * pose error - kd * relative site velocity (osc.py:418-447, control_utils.py:85-111).
// ref_*: eef site pose (world), org_*: controller-origin site pose, goal in the origin frame, vel / bvel: site velocities
// [linear; angular] of the eef site and of the origin site.
template <typename T>
OSC_HD void osc_wrench(const T* ref_pos, const T* ref_ori, const T* org_pos, const T* org_ori, const T* goal_pos, const T* goal_ori,
                       const T* vel, const T* bvel, const double* kp, const double* kd, double* F) {
  T des_pos[3], des_ori[9], err[6], e3[3] = {0, 0, 0};
  for (int i = 0; i < 3; i++)
    des_pos[i] = org_ori[3 * i] * goal_pos[0] + org_ori[3 * i + 1] * goal_pos[1] + org_ori[3 * i + 2] * goal_pos[2] + org_pos[i];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) des_ori[3 * i + j] = org_ori[3 * i] * goal_ori[j] + org_ori[3 * i + 1] * goal_ori[3 + j] + org_ori[3 * i + 2] * goal_ori[6 + j];
  for (int k = 0; k < 3; k++) err[k] = des_pos[k] - ref_pos[k];
  for (int col = 0; col < 3; col++) {
    T rc[3] = {ref_ori[col], ref_ori[3 + col], ref_ori[6 + col]}, rd[3] = {des_ori[col], des_ori[3 + col], des_ori[6 + col]};
    e3[0] += rc[1] * rd[2] - rc[2] * rd[1];
    e3[1] += rc[2] * rd[0] - rc[0] * rd[2];
    e3[2] += rc[0] * rd[1] - rc[1] * rd[0];
  }
  for (int k = 0; k < 3; k++) err[3 + k] = T(0.5) * e3[k];
  for (int k = 0; k < 6; k++) F[k] = (double)err[k] * kp[k] - ((double)vel[k] - (double)bvel[k]) * kd[k];
}
