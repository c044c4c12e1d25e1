// Unit-queue mode (b2s_set_mode(h, 2)): the whole control step of every environment as ONE persistent kernel.
//
// The phase pipeline (b2s_pipeline.cuh) runs a substep of an environment group as three kernels; every kernel lasts as long as its
// slowest environment (a deep EPA, a 9-iteration Newton solve), so a group's chain costs max(P0) + max(narrow) + max(tail) per substep
// although the mean environment needs a quarter of that, and the GPU idles in the bubbles.  Here the schedulable unit is ONE SUBSTEP OF
// ONE ENVIRONMENT: resident warps pull units from a ticket ring in global memory, run kinematics / dynamics / broad phase, the
// environment's own narrow phase, constraint rows, controller, Newton solve, integration (the very device functions of the other two
// modes, with their per-phase shared-memory layouts placed in the warp's one workspace area), and push the environment back for its
// next substep.  Nothing waits for anybody else's slow item: an expensive environment delays only itself, and the ring hands the next
// ready environment to whichever warp is free (FIFO, so all environments advance at the same rate).
//
//   ring[t], t in [0, n_env * nsub): ticket t's environment, encoded env + n_env * substep; -1 = not produced yet.  Consumers take
//     tickets with atomicAdd(head) and wait for their slot; a finished unit with substeps left publishes the environment at
//     atomicAdd(tail).  One slot per ticket of the control step: no reuse, no ABA.
//   Environments whose contacts / constraint rows do not fit the small-tier layout go to a second ring served by the large-role
//     warps of blocks [0, n_large) (fewer warps per block, the full-capacity layout), exactly the two-tier scheme of the pipeline.
//   Memory ordering: a unit's state round-trips through global memory; the hand-over is st.release.gpu / ld.acquire.gpu on the ring
//     slot (plus a proxy fence in front of TMA reads of rows another warp wrote through the async proxy).
#pragma once
#include "b2s_pipeline.cuh"

struct UnitQ {
  int* ring;      // [total]
  int* ovf_ring;  // [total]
  int* ctr;       // [8]: 0 head ticket, 1 tail ticket, 2 finished units, 3 overflow head, 4 overflow tail
  int total, n_large, wpb_large;
  int stride, stride_large;  // words of shared memory per warp: small role / large role
  unsigned long long* prof;  // B2S_UNIT_PROF: [16] clock64 cycles per stage summed over the blocks' rounds (thread 0 of every block), [15] = rounds
  int barriers;              // small role: -1 free-running warps, 0 the block starts its round of units together, 1..4 block
                             // barriers between the stages of a round as well (EXPERIMENTAL: stalls)
};

DEV int ld_acquire_gpu(const int* p) { int v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
DEV int ld_relaxed_gpu(const int* p) { int v; asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
DEV void st_release_gpu(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
DEV void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

template <typename R> __global__ void unit_init_kernel(UnitQ q, int n_env) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < q.total) { q.ring[i] = i < n_env ? i : -1; q.ovf_ring[i] = -1; }
  if (i < 8) q.ctr[i] = i == 1 ? n_env : 0;
}

// after the persistent kernel: a watchdog event (ctr[7] != 0: a ticket never arrived, the blocks drained) means the control step is
// INCOMPLETE - flag every environment (warn bit 64) so that the caller sees it in info["sim_warn"] without a host sync
template <typename R> __global__ void unit_check_kernel(UnitQ q, int slot) {
  const DState<R>& s = cstate<R>(slot);
  if (q.ctr[7] == 0) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < s.n_env; i += gridDim.x * blockDim.x) s.warn[i] |= 64;
}

// steps 1: kinematics, velocity stage + RNE bias, CRB -> M, broad phase; candidate table of the environment; poses etc. -> workspace row
template <typename R> DEVN int unit_phase0(R* area, int lane, int slot, int env) {
  const DModel<R>& m = cmodel<R>(slot);
  const DState<R>& s = cstate<R>(slot);
  const WSLayout& L = c_lay[slot][LAY_P0];
  const WSLayout& RL = c_lay[slot][LAY_ROW];
  Eng<R> e(area, lane, slot, LAY_P0);
  e.env = env;
  size_t E = env;
  R* row = s.wsg + E * RL.total;
  load_row(e.p(L.qpos), s.qpos + E * m.nq, m.nq, lane);
  load_row(e.p(L.qvel), s.qvel + E * m.nv, m.nv, lane);
  __syncwarp();
  const int was_reset = e.kinematics();
  if (was_reset) {  // diverged state reset to the model defaults (mj_checkPos / mj_checkVel): the tail reads the state from global memory
    for (int i = lane; i < m.nq; i += 32) s.qpos[E * m.nq + i] = e.p(L.qpos)[i];
    for (int i = lane; i < m.nv; i += 32) { s.qvel[E * m.nv + i] = 0; s.qacc[E * m.nv + i] = 0; s.qacc_ws[E * m.nv + i] = 0; }
    if (lane == 0) s.time[env] = 0;
    __syncwarp();
  }
  e.velocity();
  e.crb();
  int* cand = reinterpret_cast<int*>(e.p(L.scratch));
  int* cand_g = cand + 96;
  int na, ng, warn = was_reset;
  cull_pairs(e, cand, cand_g, s.cl_maxa, s.cl_maxg, na, ng);
  if (na > s.cl_maxa) { na = s.cl_maxa; warn |= 4; }
  if (ng > s.cl_maxg) { ng = s.cl_maxg; warn |= 4; }
  // the environment owns fixed output slots: analytic candidate i -> env * cl_maxa + i, convex candidate i -> env * cl_maxg + i
  int* tab = s.cl_env + E * CL_ENVW(s);
  if (lane == 0) { tab[0] = na; tab[1] = ng; }
  for (int i = lane; i < na; i += 32) { tab[2 + 2 * i] = cand[i]; tab[3 + 2 * i] = env * s.cl_maxa + i; }
  for (int i = lane; i < ng; i += 32) { tab[2 + 2 * (s.cl_maxa + i)] = cand_g[i]; tab[3 + 2 * (s.cl_maxa + i)] = env * s.cl_maxg + i; }
  if (lane == 0) reinterpret_cast<int*>(row + RL.hdr)[2] = warn;
  __syncwarp();
  ws_store(e, row, c_pio[slot][PIO_P0]);
  return na | (ng << 16);
}

// narrow phase of ONE environment by its own warp: analytic pairs one per lane, convex pairs one after the other with the warp's whole
// workspace area as EPA polytope + vertex staging scratch (phase 0's regions are in the global row by now)
template <typename R> DEVN void unit_narrow(R* area, int area_words, int lane, int slot, int env, int na, int ng) {
  const DModel<R>& m = cmodel<R>(slot);
  const DState<R>& s = cstate<R>(slot);
  const WSLayout& RL = c_lay[slot][LAY_ROW];
  size_t E = env;
  const R* row = s.wsg + E * RL.total;
  const int* tab = s.cl_env + E * CL_ENVW(s);
  for (int base = 0; base < na; base += 32) {
    int i = base + lane;
    if (i < na) {
      int pidx = tab[2 + 2 * i];
      int g1 = m.pair_geom[2 * pidx], g2 = m.pair_geom[2 * pidx + 1];
      if (m.geom_type[g1] > m.geom_type[g2]) { int t = g1; g1 = g2; g2 = t; }
      Shape<R> A, B;
      shape_from(m, g1, row + RL.gpos, row + RL.gmat, A);
      shape_from(m, g2, row + RL.gpos, row + RL.gmat, B);
      R buf[8 * CREC];
      int n = narrow_analytic(A, B, buf);
      R* out = s.cl_outA + ((size_t)env * s.cl_maxa + i) * CL_RECA;
      out[0] = R(n);
      for (int k = 0; k < n * CREC; k++) out[1 + k] = buf[k];
    }
  }
  __syncwarp();
  const int stage_cap = area_words - EPA_PIPE_WORDS;
  for (int i = 0; i < ng; i++) {
    int pidx = tab[2 + 2 * (s.cl_maxa + i)];
    int g1 = m.pair_geom[2 * pidx], g2 = m.pair_geom[2 * pidx + 1];
    if (m.geom_type[g1] > m.geom_type[g2]) { int t = g1; g1 = g2; g2 = t; }
    Shape<R> A, B;
    shape_from(m, g1, row + RL.gpos, row + RL.gmat, A);
    shape_from(m, g2, row + RL.gpos, row + RL.gmat, B);
    R buf[CREC];
    int n = convex_convex(A, B, buf, 1, area, lane, s.gjk_cache ? s.gjk_cache + ((size_t)env * m.npair + pidx) * 3 : (R*)nullptr,
                          EPA_PIPE_MAXV, EPA_PIPE_MAXF, stage_cap >= 64 ? area + EPA_PIPE_WORDS : (R*)nullptr, stage_cap >= 64 ? stage_cap : 0);
    R* out = s.cl_outG + ((size_t)env * s.cl_maxg + i) * 8;
    if (lane == 0) {
      out[0] = R(n);
      for (int k = 0; k < CREC; k++) out[1 + k] = n ? buf[k] : R(0);
    }
    __syncwarp();
  }
  __syncwarp();
}

// contact gather, constraint rows, controller, actuation, Newton solve, Euler, observation / task rows of substep `sub`.
// Returns 0 when the unit is finished, 1 when the environment does not fit this tier (nothing of its state has been touched).
template <typename R>
DEV int unit_tail(R* area, int lane, int slot, int lid, int env, int sub, int nsub, int phases, const R* action, unsigned long long* bar, unsigned& parity) {
  const DModel<R>& m = cmodel<R>(slot);
  const DState<R>& s = cstate<R>(slot);
  const WSLayout& L = c_lay[slot][lid];
  const WSLayout& RL = c_lay[slot][LAY_ROW];
  const PhaseIO& io = c_pio[slot][lid == LAY_TL ? PIO_TL : PIO_TS];
  const PhaseIO& io_late = c_pio[slot][lid == LAY_TL ? PIO_TL_LATE : PIO_TS_LATE];
  const CtrlCfgDev& cc = c_cc[slot];
  Eng<R> e(area, lane, slot, lid);
  const bool tiered = L.mc < m.maxcon || L.me < m.maxefc;
  const size_t E = env;
  const R* row = s.wsg + E * RL.total;
  int warn = reinterpret_cast<const int*>(row + RL.hdr)[2];
  ws_load(e, row, io, bar, parity);
  load_row(e.p(L.qpos), s.qpos + E * m.nq, m.nq, lane);
  load_row(e.p(L.qvel), s.qvel + E * m.nv, m.nv, lane);
  load_row(e.p(L.ctrl), s.ctrl + E * m.nu, m.nu, lane);
  load_row(e.p(L.qacc_ws), s.qacc_ws + E * m.nv, m.nv, lane);
  __syncwarp();
  int wl = 0;
  int ncon = gather_contacts(e, env, wl);
  int nefc = (tiered && (wl & 4)) ? 0 : make_constraint(e, ncon, wl);
  wl = warp_or_i(wl);
  if (tiered && (wl & 12)) return 1;
  warn |= wl;
  if (phases & PH_CTRL) {
    CtrlState<R> cs;
    ctrl_load(e, cs, env);
    ctrl_run(e, cs, env, sub == 0 ? action : (const R*)nullptr);
    for (int i = lane; i < m.nu; i += 32) s.ctrl[E * m.nu + i] = e.p(L.ctrl)[i];
    if (sub == 0) ctrl_store(e, cs, env);
    __syncwarp();
  }
  R time = s.time[env];
  e.actuation((R*)nullptr);
  if (e.acceleration()) warn |= 1;
  solve(e, nefc, ncon, warn);
  if (!(phases & PH_NOINTEGRATE)) {
    { int eb = e.euler(&time); if (eb & 32) warn |= 32; else if (eb) warn |= 2; }
  }
  if ((phases & PH_OBS) && cc.obs_dim > 0 && sub == nsub - 1) {
    ws_load(e, row, io_late, bar, parity);
    write_obs(e, env, (phases & PH_NOINTEGRATE) != 0);
    write_task(e, env, ncon);
  }
  for (int i = lane; i < m.nq; i += 32) s.qpos[E * m.nq + i] = e.p(L.qpos)[i];
  for (int i = lane; i < m.nv; i += 32) {
    s.qvel[E * m.nv + i] = e.p(L.qvel)[i];
    s.qacc[E * m.nv + i] = e.p(L.qacc)[i];
    s.qacc_ws[E * m.nv + i] = e.p(L.qacc_ws)[i];
  }
  warn = warp_or_i(warn);
  if (lane == 0) { s.time[env] = time; s.warn[env] |= warn; }
  __syncwarp();
  return 0;
}

// ---- the tail of a unit as separately compiled (noinline) stages: the lockstep rounds of the small role put block barriers between
// them, and a kernel body that inlines all of it is the kind of function nvcc 12.9 has mis-allocated before
// packed result of stage A: ncon (8 bits) | nefc (10 bits) << 8 | warn (8 bits) << 20 | does-not-fit-this-tier << 30
template <typename R>
DEVN int unit_tail_a(R* area, int lane, int slot, int env, unsigned long long* bar, unsigned& parity) {
  const DModel<R>& m = cmodel<R>(slot);
  const DState<R>& s = cstate<R>(slot);
  const WSLayout& L = c_lay[slot][LAY_TS];
  const WSLayout& RL = c_lay[slot][LAY_ROW];
  Eng<R> e(area, lane, slot, LAY_TS);
  const bool tiered = L.mc < m.maxcon || L.me < m.maxefc;
  const size_t E = env;
  const R* row = s.wsg + E * RL.total;
  int warn = reinterpret_cast<const int*>(row + RL.hdr)[2];
  ws_load(e, row, c_pio[slot][PIO_TS], bar, parity);
  load_row(e.p(L.qpos), s.qpos + E * m.nq, m.nq, lane);
  load_row(e.p(L.qvel), s.qvel + E * m.nv, m.nv, lane);
  load_row(e.p(L.ctrl), s.ctrl + E * m.nu, m.nu, lane);
  load_row(e.p(L.qacc_ws), s.qacc_ws + E * m.nv, m.nv, lane);
  __syncwarp();
  int wl = 0;
  int ncon = gather_contacts(e, env, wl);
  int nefc = (tiered && (wl & 4)) ? 0 : make_constraint(e, ncon, wl);
  wl = warp_or_i(wl);
  int ovf = (tiered && (wl & 12)) ? 1 : 0;
  warn = warp_or_i(warn | wl) & 255;
  return (ncon & 255) | ((nefc & 1023) << 8) | (warn << 20) | (ovf << 30);
}
template <typename R> DEVN void unit_tail_ctrl(R* area, int lane, int slot, int env, int sub, const R* action) {
  const DModel<R>& m = cmodel<R>(slot);
  const DState<R>& s = cstate<R>(slot);
  const WSLayout& L = c_lay[slot][LAY_TS];
  Eng<R> e(area, lane, slot, LAY_TS);
  const size_t E = env;
  CtrlState<R> cs;
  ctrl_load(e, cs, env);
  ctrl_run(e, cs, env, sub == 0 ? action : (const R*)nullptr);
  for (int i = lane; i < m.nu; i += 32) s.ctrl[E * m.nu + i] = e.p(L.ctrl)[i];
  if (sub == 0) ctrl_store(e, cs, env);
  __syncwarp();
}
template <typename R> DEVN int unit_tail_acc(R* area, int lane, int slot) {
  Eng<R> e(area, lane, slot, LAY_TS);
  e.actuation((R*)nullptr);
  int w = e.acceleration() ? 1 : 0;
  return warp_or_i(w);
}
template <typename R> DEVN int unit_tail_solve(R* area, int lane, int slot, int nefc, int ncon) {
  Eng<R> e(area, lane, slot, LAY_TS);
  int warn = 0;
  solve(e, nefc, ncon, warn);
  return warp_or_i(warn);
}
template <typename R>
DEVN void unit_tail_end(R* area, int lane, int slot, int env, int sub, int nsub, int phases, int ncon, int warn, unsigned long long* bar, unsigned& parity) {
  const DModel<R>& m = cmodel<R>(slot);
  const DState<R>& s = cstate<R>(slot);
  const WSLayout& L = c_lay[slot][LAY_TS];
  const WSLayout& RL = c_lay[slot][LAY_ROW];
  const CtrlCfgDev& cc = c_cc[slot];
  Eng<R> e(area, lane, slot, LAY_TS);
  const size_t E = env;
  const R* row = s.wsg + E * RL.total;
  R time = s.time[env];
  if (!(phases & PH_NOINTEGRATE)) {
    { int eb = e.euler(&time); if (eb & 32) warn |= 32; else if (eb) warn |= 2; }
  }
  if ((phases & PH_OBS) && cc.obs_dim > 0 && sub == nsub - 1) {
    ws_load(e, row, c_pio[slot][PIO_TS_LATE], bar, parity);
    write_obs(e, env, (phases & PH_NOINTEGRATE) != 0);
    write_task(e, env, ncon);
  }
  for (int i = lane; i < m.nq; i += 32) s.qpos[E * m.nq + i] = e.p(L.qpos)[i];
  for (int i = lane; i < m.nv; i += 32) {
    s.qvel[E * m.nv + i] = e.p(L.qvel)[i];
    s.qacc[E * m.nv + i] = e.p(L.qacc)[i];
    s.qacc_ws[E * m.nv + i] = e.p(L.qacc_ws)[i];
  }
  warn = warp_or_i(warn);
  if (lane == 0) { s.time[env] = time; s.warn[env] |= warn; }
  __syncwarp();
}

// the unit is finished: hand the environment to whoever takes the next ticket (or count it as done after its last substep)
DEV void unit_finish(const UnitQ& q, int n_env, int env, int sub, int nsub, int lane) {
  __syncwarp();
  if (lane == 0) {
    __threadfence();
    if (sub + 1 < nsub) {
      int p = atomicAdd(q.ctr + 1, 1);
      st_release_gpu(q.ring + p, env + n_env * (sub + 1));
    }
    atomicAdd(q.ctr + 2, 1);
  }
  __syncwarp();
}


// One block of 16 warps per SM: what counts is that an SM executes ONE code region at a time (the hot code is ~10x the instruction
// cache).  Measured on 4096 Lift environments (tools/run30.sh): 16 warps x 1 block 306 k env-steps/s, 8 warps x 2 blocks 248 k,
// 8 warps x 1 block 210 k, 4 warps x 4 blocks 158 k, free-running warps 80 k.
#ifndef B2S_LBU_THREADS
#define B2S_LBU_THREADS 512
#define B2S_LBU_BLOCKS 1
#endif

template <typename R>
__global__ void __launch_bounds__(B2S_LBU_THREADS, B2S_LBU_BLOCKS) unit_kernel(int phases, int nsub, const R* action, int slot, UnitQ q) {
  const DState<R>& s = cstate<R>(slot);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  R* smem = reinterpret_cast<R*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_env = s.n_env;
  __shared__ unsigned long long mbar[32];
  if (lane == 0) mbar_init(&mbar[warp]);
  __syncwarp();
  unsigned parity = 0;
  if ((int)blockIdx.x < q.n_large) {
    // ---- large role: environments the small tier could not hold (phase 0 and the narrow phase are done; their results are in the row)
    if (warp >= q.wpb_large) return;
    R* area = smem + (size_t)warp * q.stride_large;
    for (;;) {
      int item = -1;
      if (lane == 0) {
        for (;;) {
          int h = ld_relaxed_gpu(q.ctr + 3), t = ld_acquire_gpu(q.ctr + 4);
          if (h < t) { if (atomicCAS(q.ctr + 3, h, h + 1) == h) { item = h; break; } continue; }
          if (ld_acquire_gpu(q.ctr + 2) >= q.total || ld_relaxed_gpu(q.ctr + 7) != 0) break;
          __nanosleep(400);
        }
      }
      item = __shfl_sync(B2S_FULL, item, 0);
      if (item < 0) break;
      int code = 0;
      if (lane == 0) { while ((code = ld_acquire_gpu(q.ovf_ring + item)) < 0) __nanosleep(100); }
      code = __shfl_sync(B2S_FULL, code, 0);
      __syncwarp();
      fence_proxy_async_all();  // the row was written through the async proxy of another SM
      const int env = code % n_env, sub = code / n_env;
      unit_tail<R>(area, lane, slot, LAY_TL, env, sub, nsub, phases, action, &mbar[warp], parity);
      unit_finish(q, n_env, env, sub, nsub, lane);
    }
    return;
  }
  // ---- small role: the block takes `wpb` consecutive tickets per round and walks them through the stages in LOCKSTEP (block barriers
  // between the stages).  Free-running warps - every warp of an SM somewhere else in 300 KB of SASS - missed the instruction cache
  // on nearly every fetch (measured: 1.1 ms per unit against ~0.2 ms of work); in lockstep an SM executes one or two code regions at
  // a time, like the phase kernels, and a stage costs the slowest of the block's 8 units (1.0-1.3x the mean) instead of the slowest
  // of a 512-environment launch (2.5x).
  const int wpb = blockDim.x >> 5;
  __shared__ int sh_t0, sh_k;
  R* area = smem + (size_t)warp * q.stride;
#define UBAR(level) if (q.barriers >= (level)) __syncthreads();
#define UTICK(k) if (q.prof != nullptr && threadIdx.x == 0) { long long tn_ = clock64(); atomicAdd(q.prof + (k), (unsigned long long)(tn_ - tprev)); tprev = tn_; }
  long long tprev = clock64();
  for (;;) {
    int t;
    if (q.barriers >= 0) {
      // the block takes up to `wpb` tickets that are ALREADY PRODUCED (head < tail).  The first lockstep version took `wpb` tickets
      // whether they existed or not and waited for the missing ones in front of the first stage barrier: every control step stalled until
      // the watchdog on the GPU (the ticket arithmetic itself terminates: tests/test_unit_queue_protocol.py).  No wait inside a round now.
      __syncthreads();
      if (threadIdx.x == 0) {
        int t0v = 0x7fffffff, k = 0, spins = 0;
        for (;;) {
          if (ld_relaxed_gpu(q.ctr + 7) != 0) break;
          const int H = ld_relaxed_gpu(q.ctr);
          if (H >= q.total) break;
          const int P = ld_acquire_gpu(q.ctr + 1);
          if (H < P) {
            k = min(wpb, P - H);
            if (atomicCAS(q.ctr, H, H + k) == H) { t0v = H; break; }
            k = 0;
            continue;
          }
          __nanosleep(100);
          if (++spins > (1 << 22)) { atomicCAS(q.ctr + 7, 0, 2); break; }
        }
        sh_t0 = t0v; sh_k = k;
      }
      __syncthreads();
      const int t0 = sh_t0;
      if (t0 == 0x7fffffff) break;
      UTICK(0)
      if (q.prof != nullptr && threadIdx.x == 0) atomicAdd(q.prof + 15, 1ull);
      t = warp < sh_k ? t0 + warp : q.total;
    } else {  // free-running warps (no block synchronisation at all): one ticket per warp
      t = 0;
      if (lane == 0) t = ld_relaxed_gpu(q.ctr + 7) != 0 ? 0x7fffffff : atomicAdd(q.ctr, 1);
      t = __shfl_sync(B2S_FULL, t, 0);
      if (t >= q.total) break;
    }
    bool live = t < q.total;
    int code = 0;
    if (live) {
      if (lane == 0) {
        // watchdog: a ticket that is not produced within ~0.5 s means the ring protocol is broken - flag it (ctr[7], with the ticket
        // and the ring counters beside it) and let every block drain instead of hanging the device
        int spins = 0;
        while ((code = ld_acquire_gpu(q.ring + t)) < 0) {
          __nanosleep(64);
          if (++spins > (1 << 22) || ((spins & 1023) == 0 && ld_relaxed_gpu(q.ctr + 7) != 0)) {
            if (atomicCAS(q.ctr + 7, 0, 1) == 0) { q.ctr[5] = t; q.ctr[6] = ld_relaxed_gpu(q.ctr + 1); }
            break;
          }
        }
      }
      code = __shfl_sync(B2S_FULL, code, 0);
      __syncwarp();
      if (code < 0) { live = false; code = 0; }
    }
    const int env = code % n_env, sub = code / n_env;
    UBAR(1)
    UTICK(1)
    int nn = 0;
    if (live) nn = unit_phase0<R>(area, lane, slot, env);
    UBAR(2)
    UTICK(2)
    if (live) unit_narrow<R>(area, q.stride, lane, slot, env, nn & 0xffff, nn >> 16);
    UBAR(1)
    UTICK(3)
    int warn = 0, ncon = 0, nefc = 0;
    if (live) {
      const int pk = unit_tail_a<R>(area, lane, slot, env, &mbar[warp], parity);
      ncon = pk & 255; nefc = (pk >> 8) & 1023; warn = (pk >> 20) & 255;
      if (pk >> 30) {  // does not fit the small tier: to the large-role warps, nothing of the state has been touched
        if (lane == 0) {
          __threadfence();
          int p = atomicAdd(q.ctr + 4, 1);
          st_release_gpu(q.ovf_ring + p, code);
        }
        __syncwarp();
        live = false;
      }
    }
    UBAR(3)
    UTICK(4)
    if (live && (phases & PH_CTRL)) unit_tail_ctrl<R>(area, lane, slot, env, sub, action);
    UBAR(3)
    UTICK(5)
    if (live) warn |= unit_tail_acc<R>(area, lane, slot);
    UBAR(4)
    UTICK(6)
    if (live) warn |= unit_tail_solve<R>(area, lane, slot, nefc, ncon);
    UBAR(4)
    UTICK(7)
    if (live) {
      unit_tail_end<R>(area, lane, slot, env, sub, nsub, phases, ncon, warn, &mbar[warp], parity);
      unit_finish(q, n_env, env, sub, nsub, lane);
    }
    UTICK(8)
  }
#undef UBAR
#undef UTICK
}
