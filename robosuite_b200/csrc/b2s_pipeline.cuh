// Pipeline mode: one substep = phase 0 (kinematics + dynamics + broad phase) | work-list narrow phase (analytic, convex) beside the
// thread-per-environment controller kernel | tail (contact gather, constraint rows, solve, integrate, observations), exchanging a
// per-environment workspace row through L2.  Same device functions as the fused kernel; what changes is scheduling and MEMORY:
// every kernel has its own compact shared-memory layout (LAY_P0 / LAY_TS / LAY_TL), so 24-28 warps are resident per SM instead of
// the 14 the one-size-fits-all layout allowed, and the tail kernel runs in two capacity tiers: the small tier holds the contact /
// row counts almost every environment has, the few that need more are re-run by the large tier (same results, no truncation).
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

// Every region of a phase's load / store list is 16-byte aligned in both its shared-memory and its global-row offset and in its
// length (build_layouts); regions adjacent on both sides are merged into spans on the host, so a phase moves its workspace with
// a handful of bulk copies: lane k issues span k.
template <typename R> DEV void ws_load(const Eng<R>& e, const R* row, const PhaseIO& io, unsigned long long* bar, unsigned& parity) {
  if (io.nload == 0) return;
#if B2S_TMA
  // the destination may have been read / written through the generic proxy before (the constraint Jacobian under the late poses, the
  // previous environment of a large-tier warp): order those accesses before the async-proxy writes of the bulk copies
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (e.lane == 0) mbar_expect_tx(bar, (unsigned)io.load_words * (unsigned)sizeof(R));
  __syncwarp();
  if (e.lane < io.nload) {
    Region r = io.load[e.lane];
    tma_load_1d(e.ws + r.off, row + r.goff, (unsigned)r.len * sizeof(R), bar);
  }
  mbar_wait(bar, parity);
  parity ^= 1u;
#else
  for (int k = 0; k < io.nload; k++) row_copy(e.ws + io.load[k].off, row + io.load[k].goff, io.load[k].len, e.lane);
  __syncwarp();
#endif
}
template <typename R> DEV void ws_store(const Eng<R>& e, R* row, const PhaseIO& io) {
#if B2S_TMA
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // this lane's generic-proxy writes -> visible to the async proxy
  __syncwarp();
  if (e.lane < io.nstore) {
    Region r = io.store[e.lane];
    tma_store_1d(row + r.goff, e.ws + r.off, (unsigned)r.len * sizeof(R));
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  __syncwarp();
#else
  for (int k = 0; k < io.nstore; k++) row_copy(row + io.store[k].goff, e.ws + io.store[k].off, io.store[k].len, e.lane);
#endif
}

// A launch covers one group of environments [env0, env0 + nenv); groups run on separate streams so that the tail of one
// group's kernel (its slowest environment) overlaps with other groups' work.  slot = descriptor slot of the owning handle.
struct Grp { int env0, nenv, gid, sub, slot; };
#define EPA_PIPE_MAXV EPA_MAXV
#define EPA_PIPE_MAXF EPA_MAXF
#define EPA_PIPE_WORDS EPA_AREA_WORDS(EPA_PIPE_MAXV, EPA_PIPE_MAXF)  // polytope area, then the vertex staging area
#define CLC(s, g) ((s).cl_cnt + 8 * (g).gid)  // this group's counters: nA, nG, overflowed envs, next convex item, next overflow item

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

// ---- phase 1: the three consumers of phase 0's outputs in ONE launch (block roles) -----------------------------------------------
// Controller, convex and analytic narrow phase are independent of each other.  As separate graph nodes on forked streams they
// serialised the environment groups (measured: any fork inside the captured graph halves the throughput); as roles of one kernel
// node they overlap with no fork: blocks [0, nG) convex narrow phase, [nG, nG + nC) controller, the rest analytic narrow phase -
// the role with the longest single work item (a deep EPA) is scheduled first.  Every block is one warp.
struct P1Cfg { int nG, nC, sub; };

// analytic pairs: ONE THREAD per candidate pair of any environment (32 different pairs per warp)
template <typename R> DEV void narrow_analytic_block(const Grp& g, int rb) {
  const DModel<R>& m = cmodel<R>(g.slot);
  const DState<R>& s = cstate<R>(g.slot);
  const WSLayout& RL = c_lay[g.slot][LAY_ROW];
  int tid = rb * 32 + threadIdx.x;
  if (tid >= CLC(s, g)[0]) return;
  tid += g.env0 * s.cl_maxa;  // this group's slice of the candidate list / output slots
  int code = s.cl_listA[tid];
  int env = code >> 12, pidx = code & 4095;
  const R* row = s.wsg + (size_t)env * RL.total;
  int g1 = m.pair_geom[2 * pidx], g2 = m.pair_geom[2 * pidx + 1];
  if (m.geom_type[g1] > m.geom_type[g2]) { int t = g1; g1 = g2; g2 = t; }
  Shape<R> A, B;
  shape_from(m, g1, row + RL.gpos, row + RL.gmat, A);
  shape_from(m, g2, row + RL.gpos, row + RL.gmat, B);
  R buf[8 * CREC];
  int n = narrow_analytic(A, B, buf);
  R* out = s.cl_outA + (size_t)tid * CL_RECA;
  out[0] = R(n);
  for (int k = 0; k < n * CREC; k++) out[1 + k] = buf[k];
}

// convex pairs: ONE WARP per candidate pair (mesh support scans split over the lanes).  The block owns one EPA polytope and the
// vertex staging area in shared memory; warps claim work items through an atomic counter, so the few expensive pairs (penetrating
// meshes: tens of EPA expansions) never hold idle neighbours resident.
template <typename R> DEV void narrow_convex_block(const Grp& g, unsigned char* smem_raw) {
  const DModel<R>& m = cmodel<R>(g.slot);
  const DState<R>& s = cstate<R>(g.slot);
  const WSLayout& RL = c_lay[g.slot][LAY_ROW];
  int lane = threadIdx.x & 31;
  R* scratch = reinterpret_cast<R*>(smem_raw);
  const int cnt = CLC(s, g)[1];
  while (true) {
    int item = 0;
    if (lane == 0) item = atomicAdd(CLC(s, g) + 3, 1);
    item = __shfl_sync(B2S_FULL, item, 0);
    if (item >= cnt) break;
    int wid = item + g.env0 * s.cl_maxg;
    int code = s.cl_listG[wid];
    int env = code >> 12, pidx = code & 4095;
    const R* row = s.wsg + (size_t)env * RL.total;
    int g1 = m.pair_geom[2 * pidx], g2 = m.pair_geom[2 * pidx + 1];
    if (m.geom_type[g1] > m.geom_type[g2]) { int t = g1; g1 = g2; g2 = t; }
    Shape<R> A, B;
    shape_from(m, g1, row + RL.gpos, row + RL.gmat, A);
    shape_from(m, g2, row + RL.gpos, row + RL.gmat, B);
    R buf[CREC];
#ifdef B2S_INSTR
    long long it0 = clock64();
#endif
    int n = convex_convex(A, B, buf, 1, scratch, lane, s.gjk_cache ? s.gjk_cache + ((size_t)env * m.npair + pidx) * 3 : (R*)nullptr,
                          EPA_PIPE_MAXV, EPA_PIPE_MAXF, m.stage_cap > 0 ? scratch + EPA_PIPE_WORDS : (R*)nullptr, m.stage_cap);
#ifdef B2S_INSTR
    if (lane == 0 && s.stats) {  // per-item cost histogram: bucket k = cycles in [2^(k+8), 2^(k+9)), by shape types (mesh-mesh / other)
      long long dt = clock64() - it0;
      int k = 0;
      while (k < 11 && (dt >> (k + 9)) > 0) k++;
      atomicAdd(s.stats + 500 - 12 * ((A.type == G_MESH && B.type == G_MESH) ? 2 : 1) + k, 1);
      if (n > 0) atomicAdd(s.stats + 18, 1);
      if (dt > (1 << 19) && s.slowlog) {  // items above 524 k cycles (~270 us): what are they?
        int slot = atomicAdd(s.stats + 20, 1);
        if (slot < 64) {
          const int* sp = reinterpret_cast<const int*>(scratch + 9 * EPA_PIPE_MAXV + 4 * EPA_PIPE_MAXF) + EPA_PIPE_MAXF + 64;
          int* o = s.slowlog + 12 * slot;
          o[0] = (int)dt; o[1] = A.type; o[2] = B.type; o[3] = A.nvert; o[4] = B.nvert; o[5] = sp[0]; o[6] = sp[1]; o[7] = sp[2]; o[8] = sp[3];
          o[9] = sp[4]; o[10] = g1; o[11] = g2;
        }
      }
    }
#endif
    R* out = s.cl_outG + (size_t)wid * 8;
    if (lane == 0) {
      out[0] = R(n);
      for (int k = 0; k < CREC; k++) out[1 + k] = n ? buf[k] : R(0);
    }
    __syncwarp();
  }
}

template <typename R>
__global__ void __launch_bounds__(32) phase1_kernel(const R* action, Grp g, P1Cfg c) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x;
#ifdef B2S_ZERO_SMEM
  for (int i = threadIdx.x; i < 6000; i += 32) reinterpret_cast<int*>(smem_raw)[i] = 0;
  __syncwarp();
#endif
#ifdef B2S_INSTR
  const DState<R>& s = cstate<R>(g.slot);
  const int kind = b < c.nG ? 2 : (b < c.nG + c.nC ? 4 : 1);
  INSTR_BEGIN(s, g, kind)
#endif
  if (b < c.nG) narrow_convex_block<R>(g, smem_raw);
  else if (b < c.nG + c.nC) ctrl_osc_block<R>(g.sub, action, g.env0, g.nenv, g.gid, g.slot, smem_raw, b - c.nG);
  else narrow_analytic_block<R>(g, b - c.nG - c.nC);
#ifdef B2S_INSTR
  INSTR_END(s, g, kind)
#endif
}

// collect this environment's contacts from the work-list outputs, in static-pair order (what the fused collide produces)
template <typename R> DEVN int gather_contacts(Eng<R> e, int env, int& warn) {
  const DModel<R>& m = e.model();
  const DState<R>& s = e.state();
  const WSLayout& L = e.lay();
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
  if (total > L.mc) { warn |= 4; if (L.mc < m.maxcon) return L.mc; }  // small tier: the caller hands the environment to the large tier
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
      if (c >= L.mc) break;
      const R* b = rec + CREC * k;
      cpos[3 * c] = b[0]; cpos[3 * c + 1] = b[1]; cpos[3 * c + 2] = b[2];
      cfr[3 * c] = b[3]; cfr[3 * c + 1] = b[4]; cfr[3 * c + 2] = b[5];
      cdist[c] = b[6];
      cint[5 * c] = g1; cint[5 * c + 1] = g2; cint[5 * c + 4] = pair;
    }
  }
  if (total > L.mc) total = L.mc;
  __syncwarp();
  finish_contacts(e, total);
  return total;
}

// Launch bounds (threads per block, resident blocks per SM the register allocation is sized for).  (256, 2) = 128 registers per
// thread: with (256, 3) = 80 registers the heavily spilling build mis-executed solve() on 21-dof models (a corrupted workspace
// pointer; compute-sanitizer: tools/run8.sh) - the same family of nvcc 12.9 stack-slot problems as DESIGN.md section 3 records.
// The kernels are latency bound at 4096 environments (every environment's warp is resident either way), so the lost occupancy
// costs nothing measurable (lb256x2 was the fastest variant of tools/run7.sh).
#ifndef B2S_LB0_THREADS
#define B2S_LB0_THREADS 256  // phase 0
#define B2S_LB0_BLOCKS 2
#endif
#ifndef B2S_LB5_THREADS
#define B2S_LB5_THREADS 256  // tail kernel
#define B2S_LB5_BLOCKS 2
#endif

// ---- phase 0: kinematics, velocity stage + RNE bias, CRB -> M, broad phase -> global candidate work lists
template <typename R>
__global__ void __launch_bounds__(B2S_LB0_THREADS, B2S_LB0_BLOCKS) phase0_kernel(int phases, Grp g) {
  const DModel<R>& m = cmodel<R>(g.slot);
  const DState<R>& s = cstate<R>(g.slot);
  const WSLayout& L = c_lay[g.slot][LAY_P0];
  const WSLayout& RL = c_lay[g.slot][LAY_ROW];
  extern __shared__ __align__(16) unsigned char smem_raw[];
  R* smem = reinterpret_cast<R*>(smem_raw);
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  int env = blockIdx.x * wpb + warp;
  INSTR_BEGIN(s, g, 0)
#ifdef B2S_INSTR
  long long instr_t0 = clock64();
#endif
  if (env >= g.nenv) return;
  env += g.env0;
  Eng<R> e(smem + (size_t)warp * L.total, lane, g.slot, LAY_P0);
  e.env = env;
#ifdef B2S_ZERO_SMEM
  for (int i = lane; i < L.total; i += 32) e.ws[i] = 0;
  __syncwarp();
#endif
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
  // collision candidates of this environment -> global work lists (slots by warp-aggregated atomics)
  int* cand = reinterpret_cast<int*>(e.p(L.scratch));
  int* cand_g = cand + 96;
  int na, ng, warn = was_reset;
  cull_pairs(e, cand, cand_g, s.cl_maxa, s.cl_maxg, na, ng);
  if (na > s.cl_maxa) { na = s.cl_maxa; warn |= 4; }
  if (ng > s.cl_maxg) { ng = s.cl_maxg; warn |= 4; }
  int baseA = 0, baseG = 0;
  if (lane == 0) {
    if (na) baseA = g.env0 * s.cl_maxa + atomicAdd(CLC(s, g), na);
    if (ng) baseG = g.env0 * s.cl_maxg + atomicAdd(CLC(s, g) + 1, ng);
  }
  baseA = __shfl_sync(B2S_FULL, baseA, 0);
  baseG = __shfl_sync(B2S_FULL, baseG, 0);
  int* tab = s.cl_env + E * CL_ENVW(s);
  if (lane == 0) { tab[0] = na; tab[1] = ng; }
  for (int i = lane; i < na; i += 32) { s.cl_listA[baseA + i] = (env << 12) | cand[i]; tab[2 + 2 * i] = cand[i]; tab[3 + 2 * i] = baseA + i; }
  for (int i = lane; i < ng; i += 32) { s.cl_listG[baseG + i] = (env << 12) | cand_g[i]; tab[2 + 2 * (s.cl_maxa + i)] = cand_g[i]; tab[3 + 2 * (s.cl_maxa + i)] = baseG + i; }
  if (lane == 0) reinterpret_cast<int*>(row + RL.hdr)[2] = warn;
  __syncwarp();
  ws_store(e, row, c_pio[g.slot][PIO_P0]);
#ifdef B2S_INSTR
  if (lane == 0 && s.cyc) s.cyc[(E * 32 + (g.sub & 31)) * 2] = (float)(clock64() - instr_t0);
#endif
  INSTR_END(s, g, 0)
}

// ---- tail: gather contacts, constraint rows + Jacobian, (in-kernel controller), actuation, Newton solve, Euler, observations.
// tier 0: warp per environment of the group, small-capacity layout; an environment whose contacts / rows do not fit is appended to
// the group's overflow list untouched.  tier 1: warps claim the overflowed environments and run them with the full-capacity layout.
template <typename R>
__global__ void __launch_bounds__(B2S_LB5_THREADS, B2S_LB5_BLOCKS) tail_kernel(int phases, int nsub, const R* action, Grp g, int tier) {
  const DModel<R>& m = cmodel<R>(g.slot);
  const DState<R>& s = cstate<R>(g.slot);
  const int lid = tier ? LAY_TL : LAY_TS;
  const WSLayout& L = c_lay[g.slot][lid];
  const WSLayout& RL = c_lay[g.slot][LAY_ROW];
  const PhaseIO& io = c_pio[g.slot][tier ? PIO_TL : PIO_TS];
  const PhaseIO& io_late = c_pio[g.slot][tier ? PIO_TL_LATE : PIO_TS_LATE];
  const CtrlCfgDev& cc = c_cc[g.slot];
  extern __shared__ __align__(16) unsigned char smem_raw[];
  R* smem = reinterpret_cast<R*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5, sub = g.sub;
  INSTR_BEGIN(s, g, tier ? 5 : 3)
  __shared__ unsigned long long mbar[32];  // one transaction barrier per warp (TMA loads of its workspace regions)
  if (lane == 0) mbar_init(&mbar[warp]);
  __syncwarp();
  unsigned parity = 0;
  Eng<R> e(smem + (size_t)warp * L.total, lane, g.slot, lid);
#ifdef B2S_ZERO_SMEM
  for (int i = lane; i < L.total; i += 32) e.ws[i] = 0;
  __syncwarp();
#endif
  int* clc = CLC(s, g);
  const bool tiered = L.mc < m.maxcon || L.me < m.maxefc;
  for (int iter = 0;; iter++) {
    int env;
    if (tier == 0) {
      if (iter > 0) break;
      env = blockIdx.x * wpb + warp;
      if (env >= g.nenv) break;
      env += g.env0;
    } else {
      int item = 0;
      if (lane == 0) item = atomicAdd(clc + 4, 1);
      item = __shfl_sync(B2S_FULL, item, 0);
      if (item >= clc[2]) break;
      env = s.ovf_list[g.env0 + item];
    }
#ifdef B2S_INSTR
    long long instr_t0 = clock64();
#endif
    const size_t E = env;
    const R* row = s.wsg + E * RL.total;
    int warn = reinterpret_cast<const int*>(row + RL.hdr)[2];  // phase 0: candidate-list overflow
    ws_load(e, row, io, &mbar[warp], parity);
    load_row(e.p(L.qpos), s.qpos + E * m.nq, m.nq, lane);
    load_row(e.p(L.qvel), s.qvel + E * m.nv, m.nv, lane);
    load_row(e.p(L.ctrl), s.ctrl + E * m.nu, m.nu, lane);
    load_row(e.p(L.qacc_ws), s.qacc_ws + E * m.nv, m.nv, lane);
    __syncwarp();
    int wl = 0;
    int ncon = gather_contacts(e, env, wl);
    int nefc = (tiered && (wl & 4)) ? 0 : make_constraint(e, ncon, wl);
    wl = warp_or_i(wl);  // make_constraint flags a dropped contact on the lane that owns it: the decision below must be warp-uniform
    if (tiered && (wl & 12)) {  // does not fit this tier: nothing of the environment's state has been touched yet
      if (lane == 0) s.ovf_list[g.env0 + atomicAdd(clc + 2, 1)] = env;
      __syncwarp();
      continue;
    }
    warn |= wl;
    if ((phases & PH_CTRL) && !(phases & PH_CTRL_EXT)) {
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
      // The reference's observables sample on the LAST substep of a control step: reset()'s forced update already
      // advances their period timer by one model timestep (utils/observables.py:214-259, environments/base.py:418-427),
      // so the period closes after substep 24 and the next update - substep 25 - takes the sample.
      // Body / site poses of this substep's step1 arrive now, over the (dead) constraint Jacobian.
      ws_load(e, row, io_late, &mbar[warp], parity);
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
#ifdef B2S_INSTR
    if (lane == 0 && s.cyc) s.cyc[(E * 32 + (sub & 31)) * 2 + 1] = (float)(clock64() - instr_t0);
    if (lane == 0 && s.stats) { atomicAdd(s.stats + 32 + min(ncon, 128), 1); atomicAdd(s.stats + 176 + min(nefc, 320), 1); if (tier) atomicAdd(s.stats + 19, 1); }
#endif
  }
  INSTR_END(s, g, tier ? 5 : 3)
}
