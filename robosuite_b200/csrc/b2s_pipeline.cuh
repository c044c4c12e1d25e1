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

#ifndef B2S_TMA
#define B2S_TMA 1  // move workspace regions with the TMA bulk-copy engine (cp.async.bulk + mbarrier)
#endif

// ---- TMA 1-D bulk copies (SASS: UBLKCP).  One elected lane per warp issues the copies; completion of loads is signalled on
// the warp's mbarrier (transaction bytes), stores are tracked with a bulk async-group.
DEV unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
DEV void mbar_init(unsigned long long* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
DEV void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
DEV void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
DEV void tma_load_1d(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
DEV void tma_store_1d(void* gdst, const void* smem_src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}

// Every region of a phase's load / store list is 16-byte aligned in offset and length (build_layout) and adjacent regions
// are merged into spans on the host, so a phase moves its workspace with a handful of bulk copies: lane k issues span k.
template <typename R> DEV void ws_load(const Eng<R>& e, const R* row, const PhaseIO& io, int nefc_nv, unsigned long long* bar) {
  const int W = 16 / (int)sizeof(R);
  int dynlen = (nefc_nv + W - 1) / W * W;
#if B2S_TMA
  if (e.lane == 0) mbar_expect_tx(bar, (unsigned)(io.load_words + (io.load_dyn ? dynlen : 0)) * (unsigned)sizeof(R));
  __syncwarp();
  if (e.lane < io.nload) {
    Region r = io.load[e.lane];
    int len = r.dyn == 1 ? dynlen : r.len;
    if (len > 0) tma_load_1d(e.ws + r.off, row + r.off, (unsigned)len * sizeof(R), bar);
  }
  mbar_wait(bar, 0);
#else
  for (int k = 0; k < io.nload; k++) row_copy(e.ws + io.load[k].off, row + io.load[k].off, io.load[k].dyn == 1 ? dynlen : io.load[k].len, e.lane);
#endif
}
template <typename R> DEV void ws_store(const Eng<R>& e, R* row, const PhaseIO& io, int nefc_nv) {
  const int W = 16 / (int)sizeof(R);
  int dynlen = (nefc_nv + W - 1) / W * W;
#if B2S_TMA
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // this lane's generic-proxy writes -> visible to the async proxy
  __syncwarp();
  if (e.lane < io.nstore) {
    Region r = io.store[e.lane];
    int len = r.dyn == 1 ? dynlen : r.len;
    if (len > 0) tma_store_1d(row + r.off, e.ws + r.off, (unsigned)len * sizeof(R));
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  __syncwarp();
#else
  for (int k = 0; k < io.nstore; k++) row_copy(row + io.store[k].off, e.ws + io.store[k].off, io.store[k].dyn == 1 ? dynlen : io.store[k].len, e.lane);
#endif
}

// A launch covers one group of environments [env0, env0 + nenv); groups run on separate streams so that the tail of one
// group's kernel (its slowest environment) overlaps with other groups' work.
struct Grp { int env0, nenv, gid, sub; };

// -DB2S_INSTR: every launch stamps its first / last %globaltimer into st_begin / st_end (device timeline of the CUDA-graph
// replay, which events cannot subdivide), warps record their clock64 cost per environment-substep.  Empty in product builds.
#ifdef B2S_INSTR
DEV unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define INSTR_SLOT(g, kind) ((((g).gid & 63) * 32 + ((g).sub & 31)) * 8 + (kind))
#define INSTR_BEGIN(st, g, kind) if (threadIdx.x == 0 && (st).st_begin) atomicMin((st).st_begin + INSTR_SLOT(g, kind), gtimer());
#define INSTR_END(st, g, kind) if ((threadIdx.x & 31) == 0 && (st).st_end) atomicMax((st).st_end + INSTR_SLOT(g, kind), gtimer());
#else
#define INSTR_BEGIN(st, g, kind)
#define INSTR_END(st, g, kind)
#endif
#include "b2s_ctrlkernel.cuh"

#define EPA_PIPE_MAXV EPA_MAXV
#define EPA_PIPE_MAXF EPA_MAXF

// ---- work-list narrow phase --------------------------------------------------------------------------------------
// analytic pairs: ONE THREAD per candidate pair of any environment (32 different pairs per warp)
template <typename R>
__global__ void __launch_bounds__(128) narrow_analytic_kernel(Grp g) {
  const DModel<R>& m = cmodel<R>();
  const DState<R>& s = cstate<R>();
  const WSLayout& L = c_L;
  int tid = blockIdx.x * blockDim.x + threadIdx.x;
  INSTR_BEGIN(s, g, 1)
  if (tid >= s.cl_cnt[4 * g.gid]) { INSTR_END(s, g, 1) return; }
  tid += g.env0 * s.cl_maxa;  // this group's slice of the candidate list / output slots
  int code = s.cl_listA[tid];
  int env = code >> 12, pidx = code & 4095;
  const R* row = s.wsg + (size_t)env * L.total;
  int g1 = m.pair_geom[2 * pidx], g2 = m.pair_geom[2 * pidx + 1];
  if (m.geom_type[g1] > m.geom_type[g2]) { int t = g1; g1 = g2; g2 = t; }
  Shape<R> A, B;
  shape_from(g1, row + L.gpos, row + L.gmat, A);
  shape_from(g2, row + L.gpos, row + L.gmat, B);
  R buf[8 * CREC];
  int n = narrow_analytic(A, B, buf);
  R* out = s.cl_outA + (size_t)tid * CL_RECA;
  out[0] = R(n);
  for (int k = 0; k < n * CREC; k++) out[1 + k] = buf[k];
  INSTR_END(s, g, 1)
}

// convex pairs: ONE WARP per candidate pair (mesh support scans split over the lanes).  A block is one warp and owns one EPA
// polytope in shared memory (7.3 KB fp32 -> ~30 resident warps per SM); warps claim work items through an atomic counter, so
// the few expensive pairs (penetrating meshes: tens of EPA expansions) never hold idle neighbours resident.
template <typename R>
__global__ void __launch_bounds__(32) narrow_convex_kernel(Grp g) {
  const DModel<R>& m = cmodel<R>();
  const DState<R>& s = cstate<R>();
  const WSLayout& L = c_L;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int lane = threadIdx.x & 31;
  R* scratch = reinterpret_cast<R*>(smem_raw);
  const int cnt = s.cl_cnt[4 * g.gid + 1];
  INSTR_BEGIN(s, g, 2)
  while (true) {
    int item = 0;
    if (lane == 0) item = atomicAdd(s.cl_cnt + 4 * g.gid + 3, 1);
    item = __shfl_sync(B2S_FULL, item, 0);
    if (item >= cnt) break;
    int wid = item + g.env0 * s.cl_maxg;
    int code = s.cl_listG[wid];
    int env = code >> 12, pidx = code & 4095;
    const R* row = s.wsg + (size_t)env * L.total;
    int g1 = m.pair_geom[2 * pidx], g2 = m.pair_geom[2 * pidx + 1];
    if (m.geom_type[g1] > m.geom_type[g2]) { int t = g1; g1 = g2; g2 = t; }
    Shape<R> A, B;
    shape_from(g1, row + L.gpos, row + L.gmat, A);
    shape_from(g2, row + L.gpos, row + L.gmat, B);
    R buf[CREC];
    int n = convex_convex(A, B, buf, 1, scratch, lane, s.gjk_cache ? s.gjk_cache + ((size_t)env * m.npair + pidx) * 3 : (R*)nullptr,
                          EPA_PIPE_MAXV, EPA_PIPE_MAXF);
    R* out = s.cl_outG + (size_t)wid * 8;
    if (lane == 0) {
      out[0] = R(n);
      for (int k = 0; k < CREC; k++) out[1 + k] = n ? buf[k] : R(0);
    }
    __syncwarp();
  }
  INSTR_END(s, g, 2)
}

// collect this environment's contacts from the work-list outputs, in static-pair order (what the fused collide produces)
template <typename R> DEVN int gather_contacts(Eng<R> e, int env, int& warn) {
  const DModel<R>& m = cmodel<R>();
  const DState<R>& s = cstate<R>();
  const WSLayout& L = c_L;
  int lane = e.lane;
  const int* tab = s.cl_env + (size_t)env * CL_ENVW(s);
  int na = tab[0], ng = tab[1], nc = na + ng;  // nc <= CL_MAXA + CL_MAXG = 96: up to IT candidates per lane
  const int IT = 3;
  int nchunk = (nc + 31) >> 5;
  int pairv[IT], slotv[IT], isgv[IT], cntv[IT], offv[IT];
#pragma unroll
  for (int it = 0; it < IT; it++) {
    int j = lane + 32 * it;
    pairv[it] = 0x7fffffff; slotv[it] = 0; isgv[it] = 0; cntv[it] = 0; offv[it] = 0;
    if (j < na) { pairv[it] = tab[2 + 2 * j]; slotv[it] = tab[3 + 2 * j]; cntv[it] = (int)s.cl_outA[(size_t)slotv[it] * CL_RECA]; }
    else if (j < nc) {
      int k = j - na;
      isgv[it] = 1; pairv[it] = tab[2 + 2 * (s.cl_maxa + k)]; slotv[it] = tab[3 + 2 * (s.cl_maxa + k)];
      cntv[it] = (int)s.cl_outG[(size_t)slotv[it] * 8];
    }
  }
  // contact offset = contacts of candidates with a smaller pair index
  int total = 0;
#pragma unroll
  for (int it2 = 0; it2 < IT; it2++) {
    if (it2 >= nchunk) break;
    int lim = nc - 32 * it2 < 32 ? nc - 32 * it2 : 32;
    for (int o = 0; o < lim; o++) {
      int op = __shfl_sync(B2S_FULL, pairv[it2], o), on = __shfl_sync(B2S_FULL, cntv[it2], o);
      total += on;
#pragma unroll
      for (int it = 0; it < IT; it++)
        if (op < pairv[it]) offv[it] += on;
    }
  }
  R* cpos = e.p(L.c_pos); R* cfr = e.p(L.c_frame); R* cdist = e.p(L.c_dist);
  int* cint = e.pi(L.c_int);
#pragma unroll
  for (int it = 0; it < IT; it++) {
    if (lane + 32 * it >= nc) continue;
    int pair = pairv[it], slot = slotv[it], n = cntv[it], off = offv[it];
    const R* rec = isgv[it] ? s.cl_outG + (size_t)slot * 8 + 1 : s.cl_outA + (size_t)slot * CL_RECA + 1;
    int g1 = m.pair_geom[2 * pair], g2 = m.pair_geom[2 * pair + 1];
    if (m.geom_type[g1] > m.geom_type[g2]) { int t = g1; g1 = g2; g2 = t; }
    for (int k = 0; k < n; k++) {
      int c = off + k;
      if (c >= m.maxcon) break;
      const R* b = rec + CREC * k;
      cpos[3 * c] = b[0]; cpos[3 * c + 1] = b[1]; cpos[3 * c + 2] = b[2];
      cfr[3 * c] = b[3]; cfr[3 * c + 1] = b[4]; cfr[3 * c + 2] = b[5];
      cdist[c] = b[6];
      cint[5 * c] = g1; cint[5 * c + 1] = g2; cint[5 * c + 4] = pair;
    }
  }
  if (total > m.maxcon) { total = m.maxcon; warn |= 4; }
  __syncwarp();
  finish_contacts(e, total);
  return total;
}

// PH: 0 kinematics+velocity+crb, 1 collision, 2 constraint rows, 3 controller, 4 actuation+solve+integrate(+obs),
//     5 = 2 + 3 + 4 in one launch (constraint rows, Jacobian and controller output never leave shared memory)
#ifndef B2S_LB_THREADS
#define B2S_LB_THREADS 512  // phase kernels: threads per block / resident blocks per SM the register allocation is sized for
#define B2S_LB_BLOCKS 1
#endif
template <typename R, int PH>
__global__ void __launch_bounds__(B2S_LB_THREADS, B2S_LB_BLOCKS) phase_kernel(int phases, int sub, int nsub, const R* action, Grp g) {
  const DModel<R>& m = cmodel<R>();
  const DState<R>& s = cstate<R>();
  const WSLayout& L = c_L;
  const PhaseIO& io = c_pio[PH];
  extern __shared__ __align__(16) unsigned char smem_raw[];
  R* smem = reinterpret_cast<R*>(smem_raw);
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  int env = blockIdx.x * wpb + warp;
  INSTR_BEGIN(s, g, PH == 0 ? 0 : 3)
#ifdef B2S_INSTR
  long long instr_t0 = clock64();
#endif
#ifndef B2S_TAIL_BARRIERS
#define B2S_TAIL_BARRIERS 0  // block barriers between the sub-phases of the merged tail kernel (instruction-cache locality)
#endif
#define TAIL_BAR(level) if (PH == 5 && B2S_TAIL_BARRIERS >= level) __syncthreads();
  if (env >= g.nenv) {  // warps past the end of the group still take part in the block barriers
    TAIL_BAR(1) TAIL_BAR(2) TAIL_BAR(3) TAIL_BAR(4)
    return;
  }
  env += g.env0;
  __shared__ unsigned long long mbar[16];  // one transaction barrier per warp (TMA loads of its workspace regions)
  if (lane == 0) mbar_init(&mbar[warp]);
  __syncwarp();
  Eng<R> e(smem + (size_t)warp * L.total, lane);
  size_t E = env;
  R* row = s.wsg + E * L.total;
  int* hdr = e.pi(L.hdr);
  if (PH >= 2) {  // ncon / nefc / warn travel in the header
    if (lane < 8) hdr[lane] = reinterpret_cast<const int*>(row + L.hdr)[lane];
    __syncwarp();
  }
  int ncon = PH >= 2 ? hdr[0] : 0, nefc = (PH == 3 || PH == 4) ? hdr[1] : 0, warn = PH >= 2 ? hdr[2] : 0;
  ws_load(e, row, io, nefc * m.nv, &mbar[warp]);
  if (PH == 0 || PH >= 2) {
    load_row(e.p(L.qpos), s.qpos + E * m.nq, m.nq, lane);
    load_row(e.p(L.qvel), s.qvel + E * m.nv, m.nv, lane);
  }
  if (PH >= 3) load_row(e.p(L.ctrl), s.ctrl + E * m.nu, m.nu, lane);
  if (PH >= 4) load_row(e.p(L.qacc_ws), s.qacc_ws + E * m.nv, m.nv, lane);
  __syncwarp();
  if (PH == 0) {
    e.kinematics();
    e.velocity();
    e.crb();
    // collision candidates of this environment -> global work lists (slots by warp-aggregated atomics)
    int* cand = reinterpret_cast<int*>(e.p(L.scratch));
    int* cand_g = cand + 96;
    int na, ng;
    cull_pairs(e, cand, cand_g, s.cl_maxa, s.cl_maxg, na, ng);
    if (na > s.cl_maxa) { na = s.cl_maxa; warn |= 4; }
    if (ng > s.cl_maxg) { ng = s.cl_maxg; warn |= 4; }
    int baseA = 0, baseG = 0;
    if (lane == 0) {
      if (na) baseA = g.env0 * s.cl_maxa + atomicAdd(s.cl_cnt + 4 * g.gid, na);
      if (ng) baseG = g.env0 * s.cl_maxg + atomicAdd(s.cl_cnt + 4 * g.gid + 1, ng);
    }
    baseA = __shfl_sync(B2S_FULL, baseA, 0);
    baseG = __shfl_sync(B2S_FULL, baseG, 0);
    int* tab = s.cl_env + E * CL_ENVW(s);
    if (lane == 0) { tab[0] = na; tab[1] = ng; }
    for (int i = lane; i < na; i += 32) { s.cl_listA[baseA + i] = (env << 12) | cand[i]; tab[2 + 2 * i] = cand[i]; tab[3 + 2 * i] = baseA + i; }
    for (int i = lane; i < ng; i += 32) { s.cl_listG[baseG + i] = (env << 12) | cand_g[i]; tab[2 + 2 * (s.cl_maxa + i)] = cand_g[i]; tab[3 + 2 * (s.cl_maxa + i)] = baseG + i; }
    hdr[0] = 0; hdr[1] = 0;
    if (lane == 0) { hdr[2] = warn; reinterpret_cast<int*>(row + L.hdr)[2] = warn; }
  } else if (PH == 1) {
    int dbgc[3] = {0, 0, 0};
    ncon = collide(e, warn, dbgc);
    if (lane == 0) { hdr[0] = ncon; hdr[1] = 0; hdr[2] = warn; hdr[3] = 0; }
    __syncwarp();
  }
  if (PH == 2 || PH == 5) {
    if (phases & PH_WORKLIST) ncon = gather_contacts(e, env, warn);
    nefc = make_constraint(e, ncon, warn);
    if (lane == 0) { hdr[0] = ncon; hdr[1] = nefc; hdr[2] = warn; }
    __syncwarp();
  }
  TAIL_BAR(1)
  if ((PH == 3 || (PH == 5 && (phases & PH_CTRL))) && !(phases & PH_CTRL_EXT)) {
    CtrlState<R> cs;
    ctrl_load(e, cs, env);
    ctrl_run(e, cs, env, sub == 0 ? action : (const R*)nullptr);
    for (int i = lane; i < m.nu; i += 32) s.ctrl[E * m.nu + i] = e.p(L.ctrl)[i];
    if (sub == 0) ctrl_store(e, cs, env);
    __syncwarp();
  }
  TAIL_BAR(2)
  if (PH == 4 || PH == 5) {
    R time = s.time[env];
    e.actuation((R*)nullptr);
    if (e.acceleration()) warn |= 1;
    TAIL_BAR(3)
    solve(e, nefc, ncon, warn);
    TAIL_BAR(4)
    if (!(phases & PH_NOINTEGRATE)) {
      if (e.euler(&time)) warn |= 2;
    }
    if ((phases & PH_OBS) && c_cc.obs_dim > 0) {
      // The reference's observables sample on the LAST substep of a control step: reset()'s forced update already
      // advances their period timer by one model timestep (utils/observables.py:214-259, environments/base.py:418-427),
      // so the period closes after substep 24 and the next update - substep 25 - takes the sample.
      if (sub == nsub - 1) { write_obs(e, env, (phases & PH_NOINTEGRATE) != 0); write_task(e, env, ncon); }
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
#ifdef B2S_INSTR
  if (lane == 0 && s.cyc && (PH == 0 || PH == 5)) s.cyc[(E * 32 + (sub & 31)) * 2 + (PH == 0 ? 0 : 1)] = (float)(clock64() - instr_t0);
  if (PH == 5 && lane == 0 && s.stats) { atomicAdd(s.stats + 16 + min(ncon, 32), 1); atomicAdd(s.stats + 64 + min(nefc, 64), 1); }
#endif
  INSTR_END(s, g, PH == 0 ? 0 : 3)
}
