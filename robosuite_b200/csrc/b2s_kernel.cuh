// Kernels: the fused per-warp step kernel (n_substeps x {step1, controller, step2} with state resident in shared
// memory), reset, and small support kernels.  See include/b2s.h for the reference calls each launch replaces.
#pragma once
#include "b2s_ctrl.cuh"

#ifndef B2S_BARRIERS
#define B2S_BARRIERS 4  // block barriers per substep that keep the warps of an SM in the same code region
#endif

template <typename R> DEV void load_row(R* dst, const R* src, int n, int lane) {
  for (int i = lane; i < n; i += 32) dst[i] = src[i];
}

template <typename R>
DEVN void export_step1(const Eng<R> e, int env, int ncon) {
  const DModel<R>& m = e.model();
  const WSLayout& L = e.lay();
  const DState<R>& s = e.state();
  int lane = e.lane;
  size_t E = env;
  load_row(s.xpos + E * 3 * m.nbody, e.p(L.xpos), 3 * m.nbody, lane);
  load_row(s.xquat + E * 4 * m.nbody, e.p(L.xquat), 4 * m.nbody, lane);
  load_row(s.xmat + E * 9 * m.nbody, e.p(L.xmat), 9 * m.nbody, lane);
  load_row(s.site_xpos + E * 3 * m.nsite, e.p(L.spos), 3 * m.nsite, lane);
  load_row(s.site_xmat + E * 9 * m.nsite, e.p(L.smat), 9 * m.nsite, lane);
  for (int k = lane; k < m.ncg; k += 32) {
    int g = m.cg_geom[k];
    for (int q = 0; q < 3; q++) s.geom_xpos[(E * m.ngeom + g) * 3 + q] = e.p(L.gpos)[3 * k + q];
    for (int q = 0; q < 9; q++) s.geom_xmat[(E * m.ngeom + g) * 9 + q] = e.p(L.gmat)[9 * k + q];
  }
  load_row(s.qM + E * m.nv * m.nv, e.p(L.M), m.nv * m.nv, lane);
  load_row(s.cdof + E * 6 * m.nv, e.p(L.cdof), 6 * m.nv, lane);
  load_row(s.qfrc_bias + E * m.nv, e.p(L.bias), m.nv, lane);
  load_row(s.qfrc_passive + E * m.nv, e.p(L.passive), m.nv, lane);
  const int* cint = e.pi(L.c_int);
  for (int c = lane; c < m.maxcon; c += 32) {
    bool v = c < ncon;
    s.contact_geom[(E * m.maxcon + c) * 2] = v ? cint[5 * c] : -1;
    s.contact_geom[(E * m.maxcon + c) * 2 + 1] = v ? cint[5 * c + 1] : -1;
    s.contact_dim[E * m.maxcon + c] = v ? cint[5 * c + 2] : 0;
    s.contact_dist[E * m.maxcon + c] = v ? e.p(L.c_dist)[c] : R(0);
    for (int q = 0; q < 3; q++) s.contact_pos[(E * m.maxcon + c) * 3 + q] = v ? e.p(L.c_pos)[3 * c + q] : R(0);
    {
      R f9[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      if (v) { f9[0] = e.p(L.c_frame)[3 * c]; f9[1] = e.p(L.c_frame)[3 * c + 1]; f9[2] = e.p(L.c_frame)[3 * c + 2]; make_frame(f9); }
      for (int q = 0; q < 9; q++) s.contact_frame[(E * m.maxcon + c) * 9 + q] = f9[q];
    }
    for (int q = 0; q < 3; q++) s.contact_friction[(E * m.maxcon + c) * 3 + q] = v ? e.p(L.c_fric)[3 * c + q] : R(0);
  }
  if (lane == 0) s.ncon[env] = ncon;
}

// constraint rows (the Jacobian overlays kinematics scratch, so this runs after make_constraint)
template <typename R>
DEVN void export_efc(const Eng<R> e, int env, int nefc) {
  const DModel<R>& m = e.model();
  const WSLayout& L = e.lay();
  const DState<R>& s = e.state();
  int lane = e.lane;
  size_t E = env;
  for (int r = lane; r < m.maxefc; r += 32) {
    bool v = r < nefc;
    s.efc_type[E * m.maxefc + r] = v ? (e.pi(L.e_int)[r] & 255) : 0;
    s.efc_aref[E * m.maxefc + r] = v ? e.p(L.e_aref)[r] : R(0);
    s.efc_D[E * m.maxefc + r] = v ? e.p(L.e_D)[r] : R(0);
    s.efc_R[E * m.maxefc + r] = v ? e.p(L.e_R)[r] : R(0);
  }
  for (int k = lane; k < nefc * m.nv; k += 32) s.efc_J[E * m.maxefc * m.nv + k] = e.p(L.J)[k];
  if (lane == 0) s.nefc[env] = nefc;
}

template <typename R>
DEVN void export_step2(const Eng<R> e, int env, int nefc, int niter) {
  const DModel<R>& m = e.model();
  const WSLayout& L = e.lay();
  const DState<R>& s = e.state();
  int lane = e.lane;
  size_t E = env;
  load_row(s.qfrc_actuator + E * m.nv, e.p(L.qact), m.nv, lane);
  load_row(s.qfrc_smooth + E * m.nv, e.p(L.qsmooth), m.nv, lane);
  load_row(s.qacc_smooth + E * m.nv, e.p(L.qaccs), m.nv, lane);
  load_row(s.qfrc_constraint + E * m.nv, e.p(L.qcon), m.nv, lane);
  for (int r = lane; r < m.maxefc; r += 32) s.efc_force[E * m.maxefc + r] = r < nefc ? e.p(L.e_force)[r] : R(0);
  if (lane == 0) s.solver_niter[env] = niter;
}

template <typename R>
__global__ void __launch_bounds__(512, 1) step_kernel(int phases, int nsub, const R* action, int slot, const uint8_t* mask = nullptr) {
  const DModel<R>& m = cmodel<R>(slot);
  const DState<R>& s = cstate<R>(slot);
  const WSLayout& L = c_lay[slot][LAY_FULL];
  extern __shared__ __align__(16) unsigned char smem_raw[];
  R* smem = reinterpret_cast<R*>(smem_raw);
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  int env = blockIdx.x * wpb + warp;
  // every warp of the block runs the phase sequence (block barriers keep the warps of an SM in the same code
  // region, which is what bounds the instruction-cache working set); warps beyond n_env shadow the last env
  bool live = env < s.n_env;
  if (!live) env = s.n_env - 1;
  if (mask) {
    // masked launch (forward pass of the environments being reset): blocks without a selected environment leave, the other warps of a
    // block shadow a selected one (they write the same values to the same addresses, like the tail warps above)
    __shared__ int pick;
    if (threadIdx.x == 0) pick = -1;
    __syncthreads();
    live = live && mask[env];
    if (live && lane == 0) atomicMax(&pick, env);
    __syncthreads();
    if (pick < 0) return;
    if (!live) env = pick;
  }
  Eng<R> e(smem + (size_t)warp * L.fused_stride, lane, slot, LAY_FULL);
  e.env = env;
  size_t E = env;
  load_row(e.p(L.qpos), s.qpos + E * m.nq, m.nq, lane);
  load_row(e.p(L.qvel), s.qvel + E * m.nv, m.nv, lane);
  load_row(e.p(L.qacc), s.qacc + E * m.nv, m.nv, lane);
  load_row(e.p(L.qacc_ws), s.qacc_ws + E * m.nv, m.nv, lane);
  load_row(e.p(L.ctrl), s.ctrl + E * m.nu, m.nu, lane);
  R time = s.time[env];
  CtrlState<R> cs;
  if (phases & PH_CTRL) ctrl_load(e, cs, env);
  __syncwarp();
  int warn = 0;
  float pc[12];
#pragma unroll
  for (int i = 0; i < 12; i++) pc[i] = 0;
  const bool prof = phases & PH_PROFILE;
  int dbgc[3] = {0, 0, 0};
  long long t0 = 0;
#define TICK(slot) if (prof) { long long t1 = clock64(); pc[slot] += (float)(t1 - t0); t0 = t1; }
#define BAR(level, slot) if (B2S_BARRIERS >= level) { __syncthreads(); TICK(slot) }
  for (int sub = 0; sub < nsub; sub++) {
    int ncon = 0, nefc = 0, niter = 0;
    bool ex = live && (phases & PH_EXPORT) && sub == nsub - 1;
    if (prof) t0 = clock64();
    BAR(1, 11)
    if (phases & PH_STEP1) {
      if (e.kinematics()) {  // diverged state reset to the model defaults (mj_checkPos / mj_checkVel)
        for (int i = lane; i < m.nv; i += 32) { e.p(L.qacc)[i] = 0; e.p(L.qacc_ws)[i] = 0; }
        time = 0; warn |= 32;
        __syncwarp();
      }
      TICK(0)
      e.velocity();
      e.crb();
      TICK(1)
      BAR(3, 11)
      ncon = collide(e, warn, dbgc, prof ? pc : (float*)nullptr);
      TICK(2)
      if (ex) export_step1(e, env, ncon);
      BAR(4, 11)
      nefc = make_constraint(e, ncon, warn);
      TICK(3)
      if (ex) export_efc(e, env, nefc);
    }
    BAR(5, 11)
    if (phases & PH_CTRL) ctrl_run(e, cs, env, sub == 0 ? action : (const R*)nullptr);
    TICK(4)
    if (phases & PH_STEP2) {
      e.actuation(ex ? s.actuator_force + E * m.nu : nullptr);
      if (e.acceleration()) warn |= 1;
      TICK(5)
      BAR(2, 11)
      niter = solve(e, nefc, ncon, warn);
      TICK(6)
      if (ex) export_step2(e, env, nefc, niter);
      BAR(6, 11)
      if (!(phases & PH_NOINTEGRATE)) {
        { int eb = e.euler(&time); if (eb & 32) warn |= 32; else if (eb) warn |= 2; }
      }
      TICK(7)
    }
    if (live && (phases & PH_OBS) && c_cc[slot].obs_dim > 0) {
      // The reference's observables sample on the LAST substep of a control step: reset()'s forced update already
      // advances their period timer by one model timestep (utils/observables.py:214-259, environments/base.py:418-427),
      // so the period closes after substep 24 and the next update - substep 25 - takes the sample.
      if (sub == nsub - 1) { write_obs(e, env, (phases & PH_NOINTEGRATE) != 0); write_task(e, env, ncon); }
    }
    __syncwarp();
  }
  if (prof && live && lane == 0)
  {
    for (int i = 0; i < 12; i++) s.prof[E * 12 + i] = pc[i];
    s.dbg[E * 4] = dbgc[0]; s.dbg[E * 4 + 1] = dbgc[1]; s.dbg[E * 4 + 2] = dbgc[2];
  }
  if (!live) return;
  // write back
  for (int i = lane; i < m.nq; i += 32) s.qpos[E * m.nq + i] = e.p(L.qpos)[i];
  for (int i = lane; i < m.nv; i += 32) {
    s.qvel[E * m.nv + i] = e.p(L.qvel)[i];
    s.qacc[E * m.nv + i] = e.p(L.qacc)[i];
    s.qacc_ws[E * m.nv + i] = e.p(L.qacc_ws)[i];
  }
  if (phases & PH_CTRL) {
    for (int i = lane; i < m.nu; i += 32) s.ctrl[E * m.nu + i] = e.p(L.ctrl)[i];
    ctrl_store(e, cs, env);
  }
  warn = warp_or_i(warn);  // some flags (a dropped contact's rows) are raised on the lane that owns the item
  if (lane == 0) { s.time[env] = time; s.warn[env] |= warn; }
}

// masked episode reset without a host round trip: selected environments take their generalized positions from `qpos_new` (a pool of
// sampled initial states, [n_env, nq]; nullptr: the model's qpos0), everything else is cleared the way reset_kernel does
template <typename R>
__global__ void reset_envs_kernel(const uint8_t* mask, const R* qpos_new, int slot) {
  const DModel<R>& m = cmodel<R>(slot);
  const DState<R>& s = cstate<R>(slot);
  int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= s.n_env) return;
  if (mask && !mask[env]) return;
  size_t E = env;
  for (int i = 0; i < m.nq; i++) s.qpos[E * m.nq + i] = qpos_new ? qpos_new[E * m.nq + i] : m.qpos0[i];
  for (int i = 0; i < m.nv; i++) { s.qvel[E * m.nv + i] = 0; s.qacc[E * m.nv + i] = 0; s.qacc_ws[E * m.nv + i] = 0; }
  for (int i = 0; i < m.nu; i++) s.ctrl[E * m.nu + i] = 0;
  s.time[env] = 0;
  s.warn[env] = 0;
  if (s.obs_fresh) s.obs_fresh[env] = 1;
}

template <typename R>
__global__ void reset_kernel(const uint8_t* mask, int slot) {
  const DModel<R>& m = cmodel<R>(slot);
  const DState<R>& s = cstate<R>(slot);
  int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= s.n_env) return;
  if (mask && !mask[env]) return;
  size_t E = env;
  for (int i = 0; i < m.nq; i++) s.qpos[E * m.nq + i] = m.qpos0[i];
  for (int i = 0; i < m.nv; i++) { s.qvel[E * m.nv + i] = 0; s.qacc[E * m.nv + i] = 0; s.qacc_ws[E * m.nv + i] = 0; }
  for (int i = 0; i < m.nu; i++) s.ctrl[E * m.nu + i] = 0;
  s.time[env] = 0;
  s.warn[env] = 0;
}

// per-episode hidden state of the collision pipeline: the GJK warm-start directions of the environments being reset
// (a replay from a restored state must not depend on what ran before: tests/test_environments/test_action_playback.py)
template <typename R>
__global__ void cache_reset_kernel(const uint8_t* mask, int slot) {
  const DModel<R>& m = cmodel<R>(slot);
  const DState<R>& s = cstate<R>(slot);
  if (!s.gjk_cache) return;
  size_t per = (size_t)m.npair * 3, idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= per * (size_t)s.n_env) return;
  if (mask && !mask[idx / per]) return;
  s.gjk_cache[idx] = 0;
}

// flattened simulator state [time, qpos, qvel] per environment <-> the state arrays (MjSimState.flatten, binding_utils.py:56-70)
template <typename R>
__global__ void state_io_kernel(R* flat, int set, int slot) {
  const DModel<R>& m = cmodel<R>(slot);
  const DState<R>& s = cstate<R>(slot);
  const int w = 1 + m.nq + m.nv;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)s.n_env * w) return;
  size_t env = idx / w; int k = (int)(idx % w);
  R* p = k == 0 ? s.time + env : (k <= m.nq ? s.qpos + env * m.nq + (k - 1) : s.qvel + env * m.nv + (k - 1 - m.nq));
  if (set) *p = flat[idx]; else flat[idx] = *p;
}

// Jacobian of a point that moves with `body` (body frame origin: kind 0, geom centre: kind 1, site: the dedicated kernel below)
template <typename R>
__global__ void jac_point_kernel(int kind, int id, R* jacp, R* jacr, int slot) {
  const DModel<R>& m = cmodel<R>(slot);
  const DState<R>& s = cstate<R>(slot);
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= s.n_env * m.nv) return;
  int env = idx / m.nv, i = idx % m.nv;
  size_t E = env;
  int body = kind == 0 ? id : m.geom_bodyid[id];
  const R* pos = kind == 0 ? s.xpos + (E * m.nbody + id) * 3 : s.geom_xpos + (E * m.ngeom + id) * 3;
  bool on = (m.body_dofmask[body] >> i) & 1ull;
  const R* cd = s.cdof + (E * m.nv + i) * 6;
  R t[3] = {0, 0, 0}, w[3] = {0, 0, 0};
  if (on) {
    v3cross(t, cd, pos);
    t[0] += cd[3]; t[1] += cd[4]; t[2] += cd[5];
    w[0] = cd[0]; w[1] = cd[1]; w[2] = cd[2];
  }
  for (int r = 0; r < 3; r++) {
    if (jacp) jacp[(E * 3 + r) * m.nv + i] = t[r];
    if (jacr) jacr[(E * 3 + r) * m.nv + i] = w[r];
  }
}

// translational / rotational Jacobian of a site from the exported cdof and site_xpos (valid after forward/step1)
template <typename R>
__global__ void jac_site_kernel(int site, R* jacp, R* jacr, int slot) {
  const DModel<R>& m = cmodel<R>(slot);
  const DState<R>& s = cstate<R>(slot);
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= s.n_env * m.nv) return;
  int env = idx / m.nv, i = idx % m.nv;
  size_t E = env;
  int body = m.site_bodyid[site];
  bool on = (m.body_dofmask[body] >> i) & 1ull;
  const R* cd = s.cdof + (E * m.nv + i) * 6;
  const R* pos = s.site_xpos + (E * m.nsite + site) * 3;
  R t[3] = {0, 0, 0}, w[3] = {0, 0, 0};
  if (on) {
    v3cross(t, cd, pos);
    t[0] += cd[3]; t[1] += cd[4]; t[2] += cd[5];
    w[0] = cd[0]; w[1] = cd[1]; w[2] = cd[2];
  }
  for (int r = 0; r < 3; r++) {
    if (jacp) jacp[(E * 3 + r) * m.nv + i] = t[r];
    if (jacr) jacr[(E * 3 + r) * m.nv + i] = w[r];
  }
}
