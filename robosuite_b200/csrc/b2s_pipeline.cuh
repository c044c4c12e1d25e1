// Pipeline mode: one substep = five small phase kernels (kinematics+dynamics | collision | constraint rows |
// controller | solve+integrate) that exchange a per-environment workspace row through L2.  Same device functions as
// the fused kernel; what changes is scheduling: every launch runs ONE phase's code on all environments, so the
// instruction working set fits the I-cache without block barriers and environments with expensive collision or many
// solver iterations no longer stall their neighbours (hardware block scheduling balances the tail).
#pragma once
#include "b2s_kernel.cuh"

template <typename R> DEV void row_copy(R* dst, const R* src, int n, int lane) {
  for (int i = lane; i < n; i += 32) dst[i] = src[i];
}
template <> DEV void row_copy<float>(float* dst, const float* src, int n, int lane) {
  // offsets and lengths of workspace regions are even: move 8 bytes per lane
  const float2* s2 = reinterpret_cast<const float2*>(src);
  float2* d2 = reinterpret_cast<float2*>(dst);
  int n2 = n >> 1;
  for (int i = lane; i < n2; i += 32) d2[i] = s2[i];
  if ((n & 1) && lane == 0) dst[n - 1] = src[n - 1];
}

template <typename R> DEV void ws_load(const Eng<R>& e, const R* row, const PhaseIO& io, int nefc_nv) {
  for (int k = 0; k < io.nload; k++) {
    int len = io.load[k].dyn == 1 ? ((nefc_nv + 1) & ~1) : io.load[k].len;
    row_copy(e.ws + io.load[k].off, row + io.load[k].off, len, e.lane);
  }
}
template <typename R> DEV void ws_store(const Eng<R>& e, R* row, const PhaseIO& io, int nefc_nv) {
  for (int k = 0; k < io.nstore; k++) {
    int len = io.store[k].dyn == 1 ? ((nefc_nv + 1) & ~1) : io.store[k].len;
    row_copy(row + io.store[k].off, e.ws + io.store[k].off, len, e.lane);
  }
}

// PH: 0 kinematics+velocity+crb, 1 collision, 2 constraint rows, 3 controller, 4 actuation+solve+integrate(+obs)
template <typename R, int PH>
__global__ void __launch_bounds__(512, 1) phase_kernel(int phases, int sub, int nsub, const R* action) {
  const DModel<R>& m = cmodel<R>();
  const DState<R>& s = cstate<R>();
  const WSLayout& L = c_L;
  const PhaseIO& io = c_pio[PH];
  extern __shared__ __align__(16) unsigned char smem_raw[];
  R* smem = reinterpret_cast<R*>(smem_raw);
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  int env = blockIdx.x * wpb + warp;
  if (env >= s.n_env) return;
  Eng<R> e(smem + (size_t)warp * L.total, lane);
  size_t E = env;
  R* row = s.wsg + E * L.total;
  int* hdr = e.pi(L.hdr);
  if (PH >= 2) {  // ncon / nefc / warn travel in the header
    if (lane < 8) hdr[lane] = reinterpret_cast<const int*>(row + L.hdr)[lane];
    __syncwarp();
  }
  int ncon = PH >= 2 ? hdr[0] : 0, nefc = PH >= 3 ? hdr[1] : 0, warn = PH >= 2 ? hdr[2] : 0;
  ws_load(e, row, io, nefc * m.nv);
  if (PH == 0 || PH == 2 || PH == 3 || PH == 4) {
    load_row(e.p(L.qpos), s.qpos + E * m.nq, m.nq, lane);
    load_row(e.p(L.qvel), s.qvel + E * m.nv, m.nv, lane);
  }
  if (PH == 3 || PH == 4) load_row(e.p(L.ctrl), s.ctrl + E * m.nu, m.nu, lane);
  if (PH == 4) load_row(e.p(L.qacc_ws), s.qacc_ws + E * m.nv, m.nv, lane);
  __syncwarp();
  if (PH == 0) {
    e.kinematics();
    e.velocity();
    e.crb();
  } else if (PH == 1) {
    int dbgc[3] = {0, 0, 0};
    ncon = collide(e, warn, dbgc);
    if (lane == 0) { hdr[0] = ncon; hdr[1] = 0; hdr[2] = warn; hdr[3] = 0; }
    __syncwarp();
  } else if (PH == 2) {
    nefc = make_constraint(e, ncon, warn);
    if (lane == 0) { hdr[1] = nefc; hdr[2] = warn; }
    __syncwarp();
  } else if (PH == 3) {
    CtrlState<R> cs;
    ctrl_load(e, cs, env);
    ctrl_run(e, cs, env, sub == 0 ? action : (const R*)nullptr);
    for (int i = lane; i < m.nu; i += 32) s.ctrl[E * m.nu + i] = e.p(L.ctrl)[i];
    if (sub == 0) ctrl_store(e, cs, env);
  } else {
    R time = s.time[env];
    e.actuation((R*)nullptr);
    if (e.acceleration()) warn |= 1;
    solve(e, nefc, ncon, warn);
    if (!(phases & PH_NOINTEGRATE)) {
      if (e.euler(&time)) warn |= 2;
    }
    if ((phases & PH_OBS) && c_cc.obs_dim > 0) {
      if (sub == 0) write_obs(e, env);
      if (sub == nsub - 1) write_task(e, env, ncon);
    }
    for (int i = lane; i < m.nq; i += 32) s.qpos[E * m.nq + i] = e.p(L.qpos)[i];
    for (int i = lane; i < m.nv; i += 32) {
      s.qvel[E * m.nv + i] = e.p(L.qvel)[i];
      s.qacc[E * m.nv + i] = e.p(L.qacc)[i];
      s.qacc_ws[E * m.nv + i] = e.p(L.qacc_ws)[i];
    }
    if (lane == 0) { s.time[env] = time; s.warn[env] |= warn; }
  }
  __syncwarp();
  ws_store(e, row, io, nefc * m.nv);
  if (PH == 1 || PH == 2) {
    if (lane < 8) reinterpret_cast<int*>(row + L.hdr)[lane] = hdr[lane];
  }
}
