// libb2s: C ABI + host side of the batched engine (model upload, workspace layout, kernel launches).
// Entry points are declared in include/b2s.h; each cites the reference call it replaces.
#include "../../include/b2s.h"
#include "b2s_unit.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <algorithm>
#include <string>
#include <vector>

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(B2S_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_)); } while (0)

// ------------------------------------------------------------------------------------------------ blob reader
struct BlobRec { char name[48]; int32_t dtype, ndim, shape[4]; int64_t off, nbytes; };
struct Blob {
  const char* p; size_t n;
  const BlobRec* find(const char* name) const {
    int64_t cnt; memcpy(&cnt, p + 8, 8);
    const BlobRec* r = (const BlobRec*)(p + 16);
    for (int64_t i = 0; i < cnt; i++) if (!strcmp(r[i].name, name)) return r + i;
    return nullptr;
  }
  bool has(const char* name) const { return find(name) != nullptr; }
  const double* f64(const char* name, int64_t* count = nullptr) const {
    const BlobRec* r = find(name);
    if (!r || r->dtype != 0) throw std::string("model blob: missing f64 field ") + name;
    if (count) *count = r->nbytes / 8;
    return (const double*)(p + r->off);
  }
  const int* i32(const char* name, int64_t* count = nullptr) const {
    const BlobRec* r = find(name);
    if (!r || r->dtype != 1) throw std::string("model blob: missing i32 field ") + name;
    if (count) *count = r->nbytes / 4;
    return (const int*)(p + r->off);
  }
  int scalar_i(const char* name) const { return i32(name)[0]; }
  double scalar_f(const char* name) const { return f64(name)[0]; }
};

// ------------------------------------------------------------------------------------------------ sim object
struct ArrayInfo { void* ptr; int dtype; int ndim; int64_t shape[4]; };

struct b2s_sim {
  int n_env = 0, device = 0, precision = B2S_F32;
  cudaStream_t stream = 0;
  std::vector<void*> allocs;
  std::map<std::string, ArrayInfo> arrays;
  WSLayout lay[B2S_NLAY]{};  // LAY_FULL (fused kernel), LAY_P0, LAY_TS / LAY_TL (tail tiers), LAY_ROW (global workspace row)
  int slot = -1;             // constant-memory descriptor slot (per device)
  int mc_small = 0, me_small = 0;  // capacities of the small tail tier (== maxcon / maxefc: no tiering)
  int osc_in_tail = 0;       // layouts built for the in-kernel OSC controller (B2S_CTRL_SPLIT=0)
  int wpb0 = 8, wpb5s = 8, wpb5l = 4;
  size_t smem0 = 0, smem5s = 0, smem5l = 0;
  DModel<float> mf{};
  DModel<double> md{};
  DState<float> sf{};
  DState<double> sd{};
  CtrlCfgDev ctrl{};
  int has_ctrl = 0;
  int wpb_fused = 4;    // fused kernel: workspace + EPA polytope area per warp
  size_t smem_fused = 0;
  int64_t launches = 0;
  int nq = 0, nv = 0, nu = 0, nbody = 0, ngeom = 0, nsite = 0, maxcon = 0, maxefc = 0, ncg = 0, hc_stride = 0;
  std::vector<double> qpos0;
  std::vector<int> site_bodyid, cgid;
  std::map<std::string, std::vector<std::string>> names;  // object type -> names by id (MjModel name tables)
  int has_obs = 0, export_env_step = 1, dirty = 1, profile = 0, mode = 0, worklist = 1, ngroups = 8;
  int timeline = 0, debug_skip = 0;  // B2S_DEBUG_SKIP: bit 0 / 1 = leave out the analytic / convex narrow-phase launch (timing experiments)
  struct TlEv { int group, type; cudaEvent_t ev; };
  std::vector<TlEv> tl_events;
  double tl_mean_us[8] = {0}; int tl_count[8] = {0};
  int ctrl_split = 1;  // pipeline: OSC controller as its own thread-per-environment kernel (B2S_CTRL_SPLIT=0: inside the tail kernel)
  std::vector<cudaStream_t> gstreams;
  std::vector<cudaEvent_t> gevents;
  cudaEvent_t fork_event = nullptr, in_event = nullptr, out_event = nullptr;
  cudaStream_t pstream = nullptr;
  void* action_buf = nullptr;
  int use_graph = 1;
  int graph_per_group = 1;  // one CUDA graph per environment group on its own stream (B2S_GRAPH_PER_GROUP=0: one graph for all)
  std::map<long long, cudaGraphExec_t> graphs;
  PhaseIO pio[B2S_NPIO];
  std::map<std::string, Region> reg;
  // unit-queue mode (mode 2, b2s_unit.cuh)
  int* uq_ring = nullptr; int* uq_ovf = nullptr; int* uq_ctr = nullptr; int uq_cap = 0;
  int uq_wpb = 0, uq_bps = 0, uq_stride = 0, uq_stride_large = 0, uq_wpb_large = 0, uq_nlarge = 0, uq_grid = 0;
  size_t uq_smem = 0;
  unsigned long long* uq_prof = nullptr;
  std::vector<double> xpos0_h, xquat0_h;  // world poses of the bodies welded to the world (model constants)
  std::vector<int> body_weldid_h;
};

template <typename T> static T* dev_upload(b2s_sim* s, const std::vector<T>& h) {
  T* d = nullptr;
  size_t n = h.size() ? h.size() : 1;
  if (cudaMalloc(&d, n * sizeof(T)) != cudaSuccess) throw std::string("cudaMalloc failed");
  s->allocs.push_back(d);
  if (h.size()) cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
  // cudaMemcpy from pageable memory returns once the data is STAGED; the DMA runs on the legacy stream, which non-blocking streams
  // (the handle's group streams, every torch stream) do not wait for: finish it here
  cudaStreamSynchronize(cudaStreamLegacy);
  return d;
}
template <typename R> static const R* up_f(b2s_sim* s, const Blob& b, const char* name) {
  int64_t n; const double* p = b.f64(name, &n);
  std::vector<R> h(n);
  for (int64_t i = 0; i < n; i++) h[i] = (R)p[i];
  return dev_upload(s, h);
}
static const int* up_i(b2s_sim* s, const Blob& b, const char* name) {
  int64_t n; const int* p = b.i32(name, &n);
  std::vector<int> h(p, p + n);
  return dev_upload(s, h);
}
template <typename R> static const R* up_vec(b2s_sim* s, const std::vector<double>& v) {
  std::vector<R> h(v.size());
  for (size_t i = 0; i < v.size(); i++) h[i] = (R)v[i];
  return dev_upload(s, h);
}

static void h_quat2mat(double* M, const double* q) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  M[0] = w * w + x * x - y * y - z * z; M[1] = 2 * (x * y - w * z); M[2] = 2 * (x * z + w * y);
  M[3] = 2 * (x * y + w * z); M[4] = w * w - x * x + y * y - z * z; M[5] = 2 * (y * z - w * x);
  M[6] = 2 * (x * z - w * y); M[7] = 2 * (y * z + w * x); M[8] = w * w - x * x - y * y + z * z;
}
static void h_qmul(double* r, const double* a, const double* b) {
  double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}

template <typename R> static void build_model(b2s_sim* s, const Blob& b, DModel<R>& m) {
  m.nq = b.scalar_i("nq"); m.nv = b.scalar_i("nv"); m.nu = b.scalar_i("nu"); m.nbody = b.scalar_i("nbody");
  m.njnt = b.scalar_i("njnt"); m.ngeom = b.scalar_i("ngeom"); m.nsite = b.scalar_i("nsite"); m.npair = b.scalar_i("npair");
  m.nmocap = b.scalar_i("nmocap");
  m.timestep = (R)b.scalar_f("opt_timestep"); m.impratio = (R)b.scalar_f("opt_impratio");
  m.density = (R)b.scalar_f("opt_density"); m.viscosity = (R)b.scalar_f("opt_viscosity");
  m.tolerance = (R)b.scalar_f("opt_tolerance"); m.meaninertia = (R)b.scalar_f("stat_meaninertia");
  m.iterations = b.scalar_i("opt_iterations"); m.ls_iterations = b.scalar_i("opt_ls_iterations");
  const double* grav = b.f64("opt_gravity");
  for (int k = 0; k < 3; k++) m.gravity[k] = (R)grav[k];
  if (b.scalar_i("opt_cone") != 1) throw std::string("only elliptic friction cones are implemented (reference models set cone=elliptic)");
  if (m.nv > 64) throw std::string("nv > 64 not supported");
  int nb = m.nbody, nv = m.nv, nj = m.njnt, ng = m.ngeom;
  const int* parent = b.i32("body_parentid");
  const int* jntnum = b.i32("body_jntnum");
  const int* jntadr = b.i32("body_jntadr");
  const int* dofnum = b.i32("body_dofnum");
  const int* dofadr = b.i32("body_dofadr");
  const int* weld = b.i32("body_weldid");
  const int* jtype = b.i32("jnt_type");
  const int* jdofadr = b.i32("jnt_dofadr");
  const int* dofjnt = b.i32("dof_jntid");
  const int* dofpar = b.i32("dof_parentid");
  const int* dofbody = b.i32("dof_bodyid");
  // one joint per body (free joints included); ball joints unsupported
  std::vector<int> body_jntid(nb, -1), depth(nb, 0), sub_end(nb, 0);
  for (int i = 0; i < nb; i++) {
    if (jntnum[i] > 1) throw std::string("bodies with more than one joint are not supported");
    if (jntnum[i] == 1) { body_jntid[i] = jntadr[i]; if (jtype[jntadr[i]] == JNT_BALL) throw std::string("ball joints are not supported"); }
  }
  int maxdepth = 0;
  for (int i = 1; i < nb; i++) {
    depth[i] = weld[i] == 0 ? 0 : depth[parent[i]] + 1;
    if (depth[i] > maxdepth) maxdepth = depth[i];
  }
  for (int i = nb - 1; i >= 0; i--) {
    if (sub_end[i] < i + 1) sub_end[i] = i + 1;
    if (i > 0 && sub_end[parent[i]] < sub_end[i]) sub_end[parent[i]] = sub_end[i];
  }
  m.maxdepth = maxdepth;
  // static world poses
  std::vector<double> xpos0(3 * nb, 0.0), xquat0(4 * nb, 0.0);
  const double* bpos = b.f64("body_pos");
  const double* bquat = b.f64("body_quat");
  xquat0[0] = 1;
  for (int i = 1; i < nb; i++) {
    if (weld[i] != 0) { xquat0[4 * i] = 1; continue; }
    double Mp[9];
    h_quat2mat(Mp, &xquat0[4 * parent[i]]);
    for (int r = 0; r < 3; r++)
      xpos0[3 * i + r] = xpos0[3 * parent[i] + r] + Mp[3 * r] * bpos[3 * i] + Mp[3 * r + 1] * bpos[3 * i + 1] + Mp[3 * r + 2] * bpos[3 * i + 2];
    h_qmul(&xquat0[4 * i], &xquat0[4 * parent[i]], bquat + 4 * i);
  }
  // dof tables
  std::vector<int> dkind(nv), cddstart(nv), fl_dof, lim_jnt;
  const double* floss = b.f64("dof_frictionloss");
  for (int i = 0; i < nv; i++) {
    int j = dofjnt[i], k = i - jdofadr[j];
    if (jtype[j] == JNT_FREE) {
      dkind[i] = k < 3 ? DK_FREE_T : DK_FREE_R;
      cddstart[i] = k < 3 ? -1 : jdofadr[j] + 2;  // rotational axes: velocity of the three translations only
    } else {
      dkind[i] = jtype[j] == JNT_SLIDE ? DK_SLIDE : DK_HINGE;
      cddstart[i] = dofpar[i];
    }
    if (floss[i] > 0) fl_dof.push_back(i);
  }
  const int* jlim = b.i32("jnt_limited");
  for (int j = 0; j < nj; j++)
    if (jlim[j] && (jtype[j] == JNT_SLIDE || jtype[j] == JNT_HINGE)) lim_jnt.push_back(j);
  m.nfl = (int)fl_dof.size(); m.nlim = (int)lim_jnt.size();
  std::vector<unsigned long long> dofmask(nb, 0ull);
  for (int i = 1; i < nb; i++) {
    dofmask[i] = dofmask[parent[i]];
    for (int d = 0; d < dofnum[i]; d++) dofmask[i] |= 1ull << (dofadr[i] + d);
  }
  // kinematic trees: dofs of one tree are contiguous (DFS numbering); world-welded bodies have tree id -1
  const int* rootid = b.i32("body_rootid");
  std::vector<int> treebase(nv, 0), treesize(nv, 0), body_tree(nb, -1);
  {
    std::vector<int> root_first(nb, -1), root_last(nb, -1);
    for (int i = 0; i < nv; i++) {
      int r = rootid[dofbody[i]];
      if (root_first[r] < 0) root_first[r] = i;
      root_last[r] = i;
    }
    int mx = 0;
    for (int i = 0; i < nv; i++) {
      int r = rootid[dofbody[i]];
      treebase[i] = root_first[r];
      treesize[i] = root_last[r] - root_first[r] + 1;
      if (treesize[i] > mx) mx = treesize[i];
    }
    m.max_treesize = mx;
    for (int i = 1; i < nb; i++) body_tree[i] = weld[i] == 0 ? -1 : rootid[i];
  }
  m.dof_treebase = dev_upload(s, treebase); m.dof_treesize = dev_upload(s, treesize); m.body_treeid = dev_upload(s, body_tree);
  std::vector<int> ment_i, ment_j;
  for (int i = 0; i < nv; i++)
    for (int j = i; j >= 0; j = dofpar[j]) { ment_i.push_back(i); ment_j.push_back(j); }
  m.nment = (int)ment_i.size();
  // colliding geoms
  const int* pair = b.i32("pair_geom");
  std::vector<int> cgid(ng, -1), cg;
  for (int p = 0; p < m.npair; p++)
    for (int k = 0; k < 2; k++) { int g = pair[2 * p + k]; if (cgid[g] < 0) cgid[g] = 1; }
  for (int g = 0; g < ng; g++) if (cgid[g] > 0) { cgid[g] = (int)cg.size(); cg.push_back(g); }
  m.ncg = (int)cg.size();
  s->cgid = cgid;
  if (m.ncg > 64) throw std::string("more than 64 colliding geoms not supported");
  const int* condim = b.i32("geom_condim");
  int maxdim = 1;
  for (int g : cg) if (condim[g] > maxdim) maxdim = condim[g];
  m.hc_stride = maxdim * maxdim;
  m.maxcon = s->maxcon; m.maxefc = s->maxefc;

  m.body_parentid = up_i(s, b, "body_parentid"); m.body_jntid = dev_upload(s, body_jntid);
  m.body_dofnum = up_i(s, b, "body_dofnum"); m.body_dofadr = up_i(s, b, "body_dofadr"); m.body_weldid = up_i(s, b, "body_weldid");
  m.body_subtree_end = dev_upload(s, sub_end); m.body_depth = dev_upload(s, depth);
  m.body_pos = up_f<R>(s, b, "body_pos"); m.body_quat = up_f<R>(s, b, "body_quat"); m.body_ipos = up_f<R>(s, b, "body_ipos");
  m.body_iquat = up_f<R>(s, b, "body_iquat"); m.body_mass = up_f<R>(s, b, "body_mass"); m.body_inertia = up_f<R>(s, b, "body_inertia");
  m.body_invweight0 = up_f<R>(s, b, "body_invweight0");
  m.body_xpos0 = up_vec<R>(s, xpos0); m.body_xquat0 = up_vec<R>(s, xquat0);
  s->xpos0_h = xpos0; s->xquat0_h = xquat0;
  { const int* wd = b.i32("body_weldid"); s->body_weldid_h.assign(wd, wd + nb); }
  m.jnt_type = up_i(s, b, "jnt_type"); m.jnt_qposadr = up_i(s, b, "jnt_qposadr"); m.jnt_dofadr = up_i(s, b, "jnt_dofadr");
  m.jnt_bodyid = up_i(s, b, "jnt_bodyid"); m.jnt_limited = up_i(s, b, "jnt_limited");
  m.jnt_pos = up_f<R>(s, b, "jnt_pos"); m.jnt_axis = up_f<R>(s, b, "jnt_axis"); m.jnt_range = up_f<R>(s, b, "jnt_range");
  m.jnt_solref = up_f<R>(s, b, "jnt_solref"); m.jnt_solimp = up_f<R>(s, b, "jnt_solimp"); m.qpos0 = up_f<R>(s, b, "qpos0");
  m.dof_bodyid = up_i(s, b, "dof_bodyid"); m.dof_jntid = up_i(s, b, "dof_jntid"); m.dof_parentid = up_i(s, b, "dof_parentid");
  m.dof_kind = dev_upload(s, dkind); m.dof_cddstart = dev_upload(s, cddstart); m.fl_dof = dev_upload(s, fl_dof);
  m.lim_jnt = dev_upload(s, lim_jnt); m.body_dofmask = dev_upload(s, dofmask);
  m.dof_armature = up_f<R>(s, b, "dof_armature"); m.dof_damping = up_f<R>(s, b, "dof_damping");
  m.dof_frictionloss = up_f<R>(s, b, "dof_frictionloss"); m.dof_solref = up_f<R>(s, b, "dof_solref");
  m.dof_solimp = up_f<R>(s, b, "dof_solimp"); m.dof_invweight0 = up_f<R>(s, b, "dof_invweight0");
  m.ment_i = dev_upload(s, ment_i); m.ment_j = dev_upload(s, ment_j);
  m.geom_type = up_i(s, b, "geom_type"); m.geom_bodyid = up_i(s, b, "geom_bodyid"); m.geom_condim = up_i(s, b, "geom_condim");
  m.geom_dataid = up_i(s, b, "geom_dataid"); m.geom_priority = up_i(s, b, "geom_priority");
  m.geom_cgid = dev_upload(s, cgid); m.cg_geom = dev_upload(s, cg);
  m.geom_size = up_f<R>(s, b, "geom_size"); m.geom_pos = up_f<R>(s, b, "geom_pos"); m.geom_quat = up_f<R>(s, b, "geom_quat");
  m.geom_friction = up_f<R>(s, b, "geom_friction"); m.geom_solmix = up_f<R>(s, b, "geom_solmix");
  m.geom_solref = up_f<R>(s, b, "geom_solref"); m.geom_solimp = up_f<R>(s, b, "geom_solimp");
  m.geom_rbound = up_f<R>(s, b, "geom_rbound"); m.geom_aabb = up_f<R>(s, b, "geom_aabb");
  m.pair_geom = up_i(s, b, "pair_geom");
  m.mesh_vertadr = up_i(s, b, "mesh_vertadr"); m.mesh_vertnum = up_i(s, b, "mesh_vertnum"); m.mesh_vert = up_f<R>(s, b, "mesh_vert");
  {  // staging area of the convex narrow-phase kernel: room for the two largest hulls (at most 56 KB), see convex_convex
    int64_t nm = 0; const int* vn = b.i32("mesh_vertnum", &nm);
    int n1 = 0, n2 = 0;
    for (int64_t i = 0; i < nm; i++) { int v = vn[i]; if (v > n1) { n2 = n1; n1 = v; } else if (v > n2) n2 = v; }
    int need = 24 + ((3 * n1 + 3) & ~3) + ((3 * n2 + 3) & ~3);  // two poses, two hulls
    int cap = (int)(56 * 1024 / sizeof(R));
    m.stage_cap = getenv("B2S_NO_STAGE") ? 0 : std::min(need, cap);
  }
  m.site_bodyid = up_i(s, b, "site_bodyid"); m.site_pos = up_f<R>(s, b, "site_pos"); m.site_quat = up_f<R>(s, b, "site_quat");
  m.act_trnid = up_i(s, b, "actuator_trnid"); m.act_ctrllimited = up_i(s, b, "actuator_ctrllimited");
  m.act_forcelimited = up_i(s, b, "actuator_forcelimited"); m.act_biastype = up_i(s, b, "actuator_biastype");
  m.act_ctrlrange = up_f<R>(s, b, "actuator_ctrlrange"); m.act_forcerange = up_f<R>(s, b, "actuator_forcerange");
  {
    int64_t n; const double* g6 = b.f64("actuator_gear", &n);
    std::vector<double> g1(m.nu);
    for (int i = 0; i < m.nu; i++) g1[i] = g6[6 * i];
    m.act_gear = up_vec<R>(s, g1);
  }
  m.act_gainprm = up_f<R>(s, b, "actuator_gainprm"); m.act_biasprm = up_f<R>(s, b, "actuator_biasprm");
}

template <typename T> static T* dev_zeros(b2s_sim* s, size_t n) {
  T* d = nullptr;
  if (n == 0) n = 1;
  if (cudaMalloc(&d, n * sizeof(T)) != cudaSuccess) throw std::string("cudaMalloc failed");
  // cudaMemset is asynchronous and runs on the legacy stream; the handle's kernels run on non-blocking streams that do not wait for
  // it (a zero-fill of the action buffer that landed in the middle of the first control step made two concurrent handles diverge)
  cudaMemset(d, 0, n * sizeof(T));
  cudaStreamSynchronize(cudaStreamLegacy);
  s->allocs.push_back(d);
  return d;
}
template <typename R> struct DT;
template <> struct DT<float> { static const int code = B2S_F32; };
template <> struct DT<double> { static const int code = B2S_F64; };

template <typename R>
static R* state_arr(b2s_sim* s, const char* name, int64_t d1, int64_t d2 = 0, int64_t d3 = 0) {
  size_t per = (size_t)(d1 ? d1 : 1) * (d2 ? d2 : 1) * (d3 ? d3 : 1);
  R* p = dev_zeros<R>(s, (size_t)s->n_env * per);
  ArrayInfo a{p, DT<R>::code, 1 + (d1 > 0) + (d2 > 0) + (d3 > 0), {s->n_env, d1, d2, d3}};
  if (d1 == 0) a.ndim = 1;
  s->arrays[name] = a;
  return p;
}
static int* state_arr_i(b2s_sim* s, const char* name, int64_t d1, int64_t d2 = 0) {
  size_t per = (size_t)(d1 ? d1 : 1) * (d2 ? d2 : 1);
  int* p = dev_zeros<int>(s, (size_t)s->n_env * per);
  ArrayInfo a{p, B2S_I32, 1 + (d1 > 0) + (d2 > 0), {s->n_env, d1, d2, 0}};
  s->arrays[name] = a;
  return p;
}

template <typename R> static void build_state(b2s_sim* s, const DModel<R>& m, DState<R>& st) {
  st.n_env = s->n_env;
  int nq = m.nq, nv = m.nv, nu = m.nu, nb = m.nbody, ng = m.ngeom, ns = m.nsite, mc = m.maxcon, me = m.maxefc;
  st.qpos = state_arr<R>(s, "qpos", nq); st.qvel = state_arr<R>(s, "qvel", nv); st.qacc = state_arr<R>(s, "qacc", nv);
  st.qacc_ws = state_arr<R>(s, "qacc_warmstart", nv); st.ctrl = state_arr<R>(s, "ctrl", nu); st.time = state_arr<R>(s, "time", 0);
  st.xpos = state_arr<R>(s, "xpos", nb, 3); st.xquat = state_arr<R>(s, "xquat", nb, 4); st.xmat = state_arr<R>(s, "xmat", nb, 9);
  st.site_xpos = state_arr<R>(s, "site_xpos", ns, 3); st.site_xmat = state_arr<R>(s, "site_xmat", ns, 9);
  st.geom_xpos = state_arr<R>(s, "geom_xpos", ng, 3); st.geom_xmat = state_arr<R>(s, "geom_xmat", ng, 9);
  st.qM = state_arr<R>(s, "qM", nv, nv); st.qfrc_bias = state_arr<R>(s, "qfrc_bias", nv);
  st.qfrc_passive = state_arr<R>(s, "qfrc_passive", nv); st.qfrc_actuator = state_arr<R>(s, "qfrc_actuator", nv);
  st.qfrc_constraint = state_arr<R>(s, "qfrc_constraint", nv); st.qfrc_smooth = state_arr<R>(s, "qfrc_smooth", nv);
  st.qacc_smooth = state_arr<R>(s, "qacc_smooth", nv); st.actuator_force = state_arr<R>(s, "actuator_force", nu);
  st.cdof = state_arr<R>(s, "cdof", nv, 6);
  st.ncon = state_arr_i(s, "ncon", 0); st.contact_geom = state_arr_i(s, "contact_geom", mc, 2);
  st.contact_dim = state_arr_i(s, "contact_dim", mc); st.nefc = state_arr_i(s, "nefc", 0);
  st.efc_type = state_arr_i(s, "efc_type", me); st.warn = state_arr_i(s, "warn", 0); st.solver_niter = state_arr_i(s, "solver_niter", 0);
  st.contact_dist = state_arr<R>(s, "contact_dist", mc); st.contact_pos = state_arr<R>(s, "contact_pos", mc, 3);
  st.contact_frame = state_arr<R>(s, "contact_frame", mc, 9); st.contact_friction = state_arr<R>(s, "contact_friction", mc, 3);
  st.contact_solref = nullptr; st.contact_solimp = nullptr;
  st.efc_J = state_arr<R>(s, "efc_J", me, nv); st.efc_force = state_arr<R>(s, "efc_force", me);
  st.efc_aref = state_arr<R>(s, "efc_aref", me); st.efc_D = state_arr<R>(s, "efc_D", me); st.efc_R = state_arr<R>(s, "efc_R", me);
  st.goal_pos = state_arr<R>(s, "ctrl_goal_pos", 3); st.goal_ori = state_arr<R>(s, "ctrl_goal_ori", 9);
  st.init_qpos_arm = state_arr<R>(s, "ctrl_initial_joint", 8); st.grip_state = state_arr<R>(s, "ctrl_grip_state", 4);
  st.ctrl_torque = state_arr<R>(s, "ctrl_torque", 8);
  st.jv_state = state_arr<R>(s, "ctrl_jv_state", 72);
  st.wsg = nullptr;
  {
    float* p = dev_zeros<float>(s, (size_t)s->n_env * 12);
    s->arrays["prof"] = ArrayInfo{p, B2S_F32, 2, {s->n_env, 12, 0, 0}};
    st.prof = p;
    st.dbg = state_arr_i(s, "dbg", 4);
  }
}

// ---- workspace layouts.  Every region is 16-byte aligned in offset and length (TMA bulk copies).
struct LayB {
  int o = 0;
  int take(int n) { int r = o; o += (n + 3) & ~3; return r; }
};
struct Dims { int nq, nv, nu, nb, ncg, ns, hc; };

// the one-size-fits-all layout of the fused kernel (every phase's regions at once)
static void layout_full(const Dims& d, int mc, int me, WSLayout& L) {
  LayB B;
  int nq = d.nq, nv = d.nv, nu = d.nu, nb = d.nb, ncg = d.ncg, ns = d.ns;
  L = WSLayout{};
  L.mc = mc; L.me = me;
  L.qpos = B.take(nq); L.qvel = B.take(nv); L.qacc = B.take(nv); L.qacc_ws = B.take(nv); L.ctrl = B.take(nu);
  L.xpos = B.take(3 * nb); L.xquat = B.take(4 * nb); L.xmat = B.take(9 * nb);
  L.cdof = B.take(6 * nv); L.cvel = B.take(6 * nb);
  L.M = B.take(nv * nv); L.H = B.take(nv * nv);
  L.bias = B.take(nv); L.passive = B.take(nv); L.qact = B.take(nv); L.qsmooth = B.take(nv); L.qaccs = B.take(nv); L.qcon = B.take(nv);
  L.spos = B.take(3 * ns); L.smat = B.take(9 * ns);
  L.c_pos = B.take(3 * mc); L.c_frame = B.take(3 * mc); L.c_dist = B.take(mc); L.c_fric = B.take(3 * mc); L.c_int = B.take(5 * mc);
  L.e_D = B.take(me); L.e_R = B.take(me); L.e_aref = B.take(me); L.e_jar = B.take(me); L.e_jv = B.take(me);
  L.e_force = B.take(me); L.e_floss = B.take(me); L.e_int = B.take(me);
  L.Ma = B.take(nv); L.grad = B.take(nv); L.search = B.take(nv); L.Mv = B.take(nv);
  // union: kinematics intermediates that are dead once collision is done  |  the constraint Jacobian
  int ubase = B.o;
  L.xipos = B.take(3 * nb); L.cdofdot = B.take(6 * nv); L.cinert = B.take(10 * nb); L.frne = B.take(6 * nb); L.ffl = B.take(6 * nb);
  L.gpos = B.take(3 * ncg); L.gmat = B.take(9 * ncg);
  L.J = ubase;
  if (ubase + me * nv > B.o) B.o = ubase + ((me * nv + 3) & ~3);
  int sc = std::max(std::max(10 * nb, 200), std::max(me + d.hc * mc + 64, 9 * mc));
  if (sc < 672) sc = 672;  // controller work area (336 doubles)
  L.scratch_size = sc; L.scratch = B.take(sc);
  L.hdr = B.take(8);
  L.total = B.o;
  L.fused_stride = L.total + EPA_AREA_WORDS(EPA_MAXV, EPA_MAXF);
}

// phase 0: kinematics, velocity stage, CRB, broad phase.  The regions it hands to the other kernels come first, in row order.
static void layout_p0(const Dims& d, WSLayout& L) {
  LayB B;
  int nq = d.nq, nv = d.nv, nb = d.nb, ncg = d.ncg, ns = d.ns;
  L = WSLayout{};
  L.qpos = B.take(nq); L.qvel = B.take(nv);
  L.xpos = B.take(3 * nb); L.xquat = B.take(4 * nb); L.cdof = B.take(6 * nv); L.cvel = B.take(6 * nb);
  bool m_own = 12 * nb < nv * nv;  // otherwise M (written by crb) lives over frne + ffl (dead after velocity)
  if (m_own) L.M = B.take(nv * nv);
  L.bias = B.take(nv); L.passive = B.take(nv); L.spos = B.take(3 * ns); L.smat = B.take(9 * ns);
  L.gpos = B.take(3 * ncg); L.gmat = B.take(9 * ncg);
  L.xmat = B.take(9 * nb); L.xipos = B.take(3 * nb); L.cdofdot = B.take(6 * nv); L.cinert = B.take(10 * nb);
  L.frne = B.take(6 * nb); L.ffl = B.take(6 * nb);
  if (!m_own) L.M = L.frne;
  int sc = std::max(10 * nb, 200);  // kinematics locals (8 nb), composite inertias (10 nb), candidate lists (96 + 32 + ...)
  L.scratch_size = sc; L.scratch = B.take(sc);
  L.hdr = B.take(8);
  L.total = B.o;
}

// tail kernel with capacities (mc, me).  osc_in_tail: the OSC controller runs inside (needs body velocities and site poses from
// the start); otherwise the poses the observation / task tables read arrive late, over the dead constraint Jacobian.
static void layout_tail(const Dims& d, int mc, int me, bool osc_in_tail, WSLayout& L) {
  LayB B;
  int nq = d.nq, nv = d.nv, nu = d.nu, nb = d.nb, ns = d.ns;
  L = WSLayout{};
  L.mc = mc; L.me = me;
  L.qpos = B.take(nq); L.qvel = B.take(nv); L.qacc = B.take(nv); L.qacc_ws = B.take(nv); L.ctrl = B.take(nu);
  L.cdof = B.take(6 * nv);
  L.M = B.take(nv * nv); L.H = B.take(nv * nv);
  L.bias = B.take(nv); L.passive = B.take(nv); L.qact = B.take(nv); L.qsmooth = B.take(nv); L.qaccs = B.take(nv); L.qcon = B.take(nv);
  L.c_pos = B.take(3 * mc); L.c_frame = B.take(3 * mc); L.c_dist = B.take(mc); L.c_fric = B.take(3 * mc); L.c_int = B.take(5 * mc);
  L.e_D = B.take(me); L.e_R = B.take(me); L.e_aref = B.take(me); L.e_jar = B.take(me); L.e_jv = B.take(me);
  L.e_force = B.take(me); L.e_floss = B.take(me); L.e_int = B.take(me);
  L.Ma = B.take(nv); L.grad = B.take(nv); L.search = B.take(nv); L.Mv = B.take(nv);
  L.J = B.o;
  int jwords = (me * nv + 3) & ~3;
  if (osc_in_tail) {
    B.o += jwords;
    L.xpos = B.take(3 * nb); L.xquat = B.take(4 * nb); L.spos = B.take(3 * ns); L.smat = B.take(9 * ns); L.cvel = B.take(6 * nb);
  } else {
    LayB O; O.o = B.o;  // overlay on J: loaded after the solve, before the observation sample
    L.xpos = O.take(3 * nb); L.xquat = O.take(4 * nb); L.spos = O.take(3 * ns); L.smat = O.take(9 * ns);
    B.o = std::max(B.o + jwords, O.o);
  }
  int sc = std::max(me + d.hc * mc + 64, 9 * mc);
  if (osc_in_tail && sc < 672) sc = 672;  // in-kernel OSC work area (336 doubles)
  L.scratch_size = sc; L.scratch = B.take(sc);
  L.hdr = B.take(8);
  L.total = B.o;
}

// global workspace row: what phase 0 hands to the narrow phase, the controller kernel and the tail
static void layout_row(const Dims& d, WSLayout& L) {
  LayB B;
  int nv = d.nv, nb = d.nb, ncg = d.ncg, ns = d.ns;
  L = WSLayout{};
  L.xpos = B.take(3 * nb); L.xquat = B.take(4 * nb); L.cdof = B.take(6 * nv); L.cvel = B.take(6 * nb);
  L.M = B.take(nv * nv); L.bias = B.take(nv); L.passive = B.take(nv); L.spos = B.take(3 * ns); L.smat = B.take(9 * ns);
  L.gpos = B.take(3 * ncg); L.gmat = B.take(9 * ncg);
  L.hdr = B.take(8);
  L.total = B.o;
}

// load / store list of one phase: (shared-memory offset, row offset, words) per named region, adjacent regions merged into spans
static void make_io(const Dims& d, const WSLayout& S, const WSLayout& ROW, std::initializer_list<const char*> names, Region* out, int& n, int* words) {
  int nv = d.nv, nb = d.nb, ncg = d.ncg, ns = d.ns;
  auto a4 = [](int x) { return (x + 3) & ~3; };
  std::vector<Region> v;
  for (const char* nm : names) {
    std::string k = nm;
    Region r{0, 0, 0, 0};
#define REG(name, field, len_) if (k == name) { r.off = S.field; r.goff = ROW.field; r.len = a4(len_); }
    REG("xpos", xpos, 3 * nb) REG("xquat", xquat, 4 * nb) REG("cdof", cdof, 6 * nv) REG("cvel", cvel, 6 * nb) REG("M", M, nv * nv)
    REG("bias", bias, nv) REG("passive", passive, nv) REG("spos", spos, 3 * ns) REG("smat", smat, 9 * ns)
    REG("gpos", gpos, 3 * ncg) REG("gmat", gmat, 9 * ncg)
#undef REG
    if (r.len > 0) v.push_back(r);
  }
  std::sort(v.begin(), v.end(), [](const Region& a, const Region& b) { return a.goff < b.goff; });
  n = 0;
  if (words) *words = 0;
  for (const Region& r : v) {
    if (n > 0 && out[n - 1].off + out[n - 1].len == r.off && out[n - 1].goff + out[n - 1].len == r.goff) out[n - 1].len += r.len;
    else { if (n >= B2S_MAXREG) throw std::string("too many workspace spans"); out[n++] = r; }
  }
  if (words) for (int k = 0; k < n; k++) *words += out[k].len;
}

static void build_layouts(b2s_sim* s, int ncg, int hc_stride) {
  Dims d{s->nq, s->nv, s->nu, s->nbody, ncg, s->nsite, hc_stride};
  layout_full(d, s->maxcon, s->maxefc, s->lay[LAY_FULL]);
  layout_p0(d, s->lay[LAY_P0]);
  layout_tail(d, s->mc_small, s->me_small, s->osc_in_tail != 0, s->lay[LAY_TS]);
  layout_tail(d, s->maxcon, s->maxefc, s->osc_in_tail != 0, s->lay[LAY_TL]);
  layout_row(d, s->lay[LAY_ROW]);
  for (int k = 0; k < B2S_NPIO; k++) s->pio[k] = PhaseIO{};
  const WSLayout& ROW = s->lay[LAY_ROW];
  make_io(d, s->lay[LAY_P0], ROW, {"xpos", "xquat", "cdof", "cvel", "M", "bias", "passive", "spos", "smat", "gpos", "gmat"},
          s->pio[PIO_P0].store, s->pio[PIO_P0].nstore, nullptr);
  for (int t = 0; t < 2; t++) {
    const WSLayout& T = s->lay[t ? LAY_TL : LAY_TS];
    PhaseIO& early = s->pio[t ? PIO_TL : PIO_TS];
    PhaseIO& late = s->pio[t ? PIO_TL_LATE : PIO_TS_LATE];
    if (s->osc_in_tail) make_io(d, T, ROW, {"cdof", "cvel", "M", "bias", "passive", "xpos", "xquat", "spos", "smat"}, early.load, early.nload, &early.load_words);
    else {
      make_io(d, T, ROW, {"cdof", "M", "bias", "passive"}, early.load, early.nload, &early.load_words);
      make_io(d, T, ROW, {"xpos", "xquat", "spos", "smat"}, late.load, late.nload, &late.load_words);
    }
  }
}

// opt a kernel in to the device's maximum dynamic shared memory (the limit is the opt-in maximum MINUS the kernel's static shared
// memory: asking for the full 227 KB on a kernel with a static mbarrier array is an invalid argument)
template <typename F> static cudaError_t optin_max_smem(F fn, int device, int* limit_out = nullptr) {
  cudaFuncAttributes a;
  cudaError_t e = cudaFuncGetAttributes(&a, fn);
  if (e != cudaSuccess) return e;
  int optin = 0;
  e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  if (e != cudaSuccess) return e;
  int lim = optin - (int)a.sharedSizeBytes;
  if (limit_out) *limit_out = lim;
  return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
}

// block shapes: warps per block bounded by the kernels' launch bounds and by 227 KB of shared memory per block
static int fit_wpb(size_t per_warp, int cap) {
  int w = cap;
  while (w > 1 && per_warp * w > 226 * 1024) w--;
  return w;
}
static int choose_blocks(b2s_sim* s) {
  size_t rsz = s->precision == B2S_F32 ? 4 : 8;
  size_t pw0 = s->lay[LAY_P0].total * rsz, pws = s->lay[LAY_TS].total * rsz, pwl = s->lay[LAY_TL].total * rsz, pwf = s->lay[LAY_FULL].fused_stride * rsz;
  if (std::max(std::max(pw0, pws), std::max(pwl, pwf)) > 226 * 1024) return fail(B2S_ERR_UNSUPPORTED, "model workspace exceeds shared memory");
  // warps per block: as many as the launch bounds allow while B2S_LBx_BLOCKS blocks still fit one SM's shared memory
  auto pick = [&](size_t pw, int lb_threads, int lb_blocks) {
    int cap = lb_threads / 32, w = cap;
    while (w > 1 && (pw * w + 1024) * lb_blocks > 228 * 1024) w--;  // 1 KB per block is reserved by the system
    if ((pw * w + 1024) * lb_blocks > 228 * 1024) w = fit_wpb(pw, cap);  // cannot reach the block count: largest block that fits
    return w;
  };
  s->wpb0 = pick(pw0, B2S_LB0_THREADS, B2S_LB0_BLOCKS);
  s->wpb5s = pick(pws, B2S_LB5_THREADS, B2S_LB5_BLOCKS);
  s->wpb5l = fit_wpb(pwl, B2S_LB5_THREADS / 32);
  if (const char* v = getenv("B2S_WPB0")) { int x = atoi(v); if (x >= 1 && x <= B2S_LB0_THREADS / 32 && pw0 * x <= 226 * 1024) s->wpb0 = x; }
  if (const char* v = getenv("B2S_WPB5")) { int x = atoi(v); if (x >= 1 && x <= B2S_LB5_THREADS / 32 && pws * x <= 226 * 1024) s->wpb5s = x; }
  s->smem0 = pw0 * s->wpb0; s->smem5s = pws * s->wpb5s; s->smem5l = pwl * s->wpb5l;
  int wf = fit_wpb(pwf, 16);
  s->wpb_fused = wf; s->smem_fused = pwf * wf;
  if (getenv("B2S_VERBOSE"))
    fprintf(stderr, "[b2s] slot %d words/warp: fused %d, P0 %d (%d warps/block), tail small %d [mc %d me %d] (%d warps/block), tail large %d [mc %d me %d] (%d), row %d\n",
            s->slot, s->lay[LAY_FULL].fused_stride, s->lay[LAY_P0].total, s->wpb0, s->lay[LAY_TS].total, s->mc_small, s->me_small, s->wpb5s,
            s->lay[LAY_TL].total, s->maxcon, s->maxefc, s->wpb5l, s->lay[LAY_ROW].total);
  return B2S_OK;
}

static b2s_sim* g_slots[64][B2S_NSLOT] = {{nullptr}};  // live handles per device: descriptor slot owners

// The phase / narrow-phase kernels need different amounts of per-thread local memory (stack).  By default the driver shrinks the
// local-memory pool after a launch and grows it again for the next kernel that needs more - a device-wide reallocation worth
// ~200 us in front of every narrow_convex launch.  Ask the context to keep the pool at its high-water mark
// (cudaDeviceLmemResizeToMax / CU_CTX_LMEM_RESIZE_TO_MAX); works on the already-initialised primary context torch created.
static void keep_local_memory_pool() {
  unsigned flags = 0;
  if (cudaGetDeviceFlags(&flags) == cudaSuccess && (flags & cudaDeviceLmemResizeToMax)) return;
  const char* how = "unchanged";
  if (cudaSetDeviceFlags((flags & cudaDeviceMask) | cudaDeviceLmemResizeToMax) == cudaSuccess) how = "cudaSetDeviceFlags";
  else {
    cudaGetLastError();
    typedef int (*get_fn)(unsigned*);
    typedef int (*set_fn)(unsigned);
    void *fg = nullptr, *fs = nullptr;
    cudaDriverEntryPointQueryResult q1, q2;
    if (cudaGetDriverEntryPoint("cuCtxGetFlags", &fg, cudaEnableDefault, &q1) == cudaSuccess &&
        cudaGetDriverEntryPoint("cuCtxSetFlags", &fs, cudaEnableDefault, &q2) == cudaSuccess && fg && fs) {
      unsigned cf = 0;
      if (((get_fn)fg)(&cf) == 0 && ((set_fn)fs)(cf | 0x10u /* CU_CTX_LMEM_RESIZE_TO_MAX */) == 0) how = "cuCtxSetFlags";
    }
    cudaGetLastError();
  }
  if (getenv("B2S_VERBOSE")) fprintf(stderr, "[b2s] local-memory pool kept at high-water mark: %s\n", how);
}


// ------------------------------------------------------------------------------------------------ API
extern "C" {

const char* b2s_last_error(void) { return g_err.c_str(); }

int b2s_create(const void* blob_host, size_t nbytes, int n_env, int device, int precision, b2s_sim** out) {
  if (!blob_host || !out || n_env <= 0) return fail(B2S_ERR_ARG, "b2s_create: bad argument");
  if (nbytes < 16 || memcmp(blob_host, "B2SMODEL", 8) != 0) return fail(B2S_ERR_MODEL, "b2s_create: not a model blob");
  if (precision != B2S_F32 && precision != B2S_F64) return fail(B2S_ERR_ARG, "b2s_create: precision must be B2S_F32 or B2S_F64");
  // work-list entries pack (env << 12 | pair) into an int, unit-queue tickets env + n_env * substep
  if (n_env >= (1 << 19)) return fail(B2S_ERR_UNSUPPORTED, "b2s_create: n_env must be below 524288 per handle (create several handles)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(B2S_ERR_CUDA, "b2s_create: no CUDA device available (this library has no CPU fallback)");
  CUDA_TRY(cudaSetDevice(device));
  if (!getenv("B2S_NO_LMEM_FLAG")) keep_local_memory_pool();
  b2s_sim* s = new b2s_sim();
  s->n_env = n_env; s->device = device; s->precision = precision;
  Blob b{(const char*)blob_host, nbytes};
  try {
    s->nq = b.scalar_i("nq"); s->nv = b.scalar_i("nv"); s->nu = b.scalar_i("nu"); s->nbody = b.scalar_i("nbody");
    s->ngeom = b.scalar_i("ngeom"); s->nsite = b.scalar_i("nsite");
    s->maxcon = b.has("opt_maxcon") ? b.scalar_i("opt_maxcon") : 32;
    s->maxefc = b.has("opt_maxefc") ? b.scalar_i("opt_maxefc") : 64;
    if (s->maxcon > 128) throw std::string("opt_maxcon > 128 not supported");
    const double* q0 = b.f64("qpos0");
    s->qpos0.assign(q0, q0 + s->nq);
    const int* sb = b.i32("site_bodyid");
    s->site_bodyid.assign(sb, sb + s->nsite);
    for (const char* ty : {"body", "joint", "geom", "site", "actuator", "mesh", "camera", "light"}) {
      std::string key = std::string("names_") + ty;
      if (!b.has(key.c_str())) continue;
      int64_t nc = 0; const int* cp = b.i32(key.c_str(), &nc);
      std::vector<std::string>& v = s->names[ty];
      std::string cur;
      for (int64_t i = 0; i < nc; i++) { if (cp[i] == '\n') { v.push_back(cur); cur.clear(); } else cur.push_back((char)cp[i]); }
      if (nc > 0) v.push_back(cur);
    }
    int ncg, hcs;
    if (precision == B2S_F32) { build_model(s, b, s->mf); build_state(s, s->mf, s->sf); ncg = s->mf.ncg; hcs = s->mf.hc_stride; }
    else { build_model(s, b, s->md); build_state(s, s->md, s->sd); ncg = s->md.ncg; hcs = s->md.hc_stride; }
    s->ncg = ncg; s->hc_stride = hcs;
    // small tail tier: capacities almost every environment of this task stays within (compiled into the model blob by the task
    // class, B2S_TIER_SMALL="mc,me" overrides); environments that need more are re-run by the large tier
    int nfl = precision == B2S_F32 ? s->mf.nfl : s->md.nfl;
    s->mc_small = b.has("opt_maxcon_small") ? b.scalar_i("opt_maxcon_small") : s->maxcon;
    s->me_small = b.has("opt_maxefc_small") ? b.scalar_i("opt_maxefc_small") : s->maxefc;
    if (const char* v = getenv("B2S_TIER_SMALL")) { int a_ = 0, b_ = 0; if (sscanf(v, "%d,%d", &a_, &b_) == 2 && a_ > 0 && b_ > 0) { s->mc_small = a_; s->me_small = b_; } }
    s->mc_small = std::min(std::max(s->mc_small, 4), s->maxcon);
    s->me_small = std::min(std::max(s->me_small, nfl + 8), s->maxefc);  // friction-loss rows are always present
    if (s->maxefc < nfl + 8) throw std::string("opt_maxefc too small for the model's friction-loss rows");
    // descriptor slot
    for (int k = 0; k < B2S_NSLOT && s->slot < 0; k++)
      if (!g_slots[device & 63][k]) { g_slots[device & 63][k] = s; s->slot = k; }
    if (s->slot < 0) throw std::string("more than 8 live handles on one device");
    build_layouts(s, ncg, hcs);
  } catch (const std::string& e) {
    b2s_destroy(s);
    return fail(B2S_ERR_MODEL, "b2s_create: " + e);
  }
  if (choose_blocks(s) != B2S_OK) { b2s_destroy(s); return B2S_ERR_UNSUPPORTED; }
  {
    // the attribute belongs to the FUNCTION, not to the handle: always opt in to the device maximum (a later handle with a
    // smaller workspace must not lower the limit of an earlier one - that broke mixed-task batches in round 1)
    cudaError_t e1;
    if (precision == B2S_F32) {
      e1 = optin_max_smem(step_kernel<float>, device);
      if (e1 == cudaSuccess) e1 = optin_max_smem(phase0_kernel<float>, device);
      if (e1 == cudaSuccess) e1 = optin_max_smem(tail_kernel<float>, device);
      if (e1 == cudaSuccess) e1 = optin_max_smem(phase1_kernel<float>, device);
    } else {
      e1 = optin_max_smem(step_kernel<double>, device);
      if (e1 == cudaSuccess) e1 = optin_max_smem(phase0_kernel<double>, device);
      if (e1 == cudaSuccess) e1 = optin_max_smem(tail_kernel<double>, device);
      if (e1 == cudaSuccess) e1 = optin_max_smem(phase1_kernel<double>, device);
    }
    if (e1 != cudaSuccess) { std::string msg = cudaGetErrorString(e1); b2s_destroy(s); return fail(B2S_ERR_CUDA, "cudaFuncSetAttribute: " + msg); }
  }
  *out = s;
  int rc = b2s_reset(s, nullptr);
  if (rc != 0) { b2s_destroy(s); *out = nullptr; return rc; }
  return B2S_OK;
}

void b2s_destroy(b2s_sim* s) {
  if (!s) return;
  if (s->uq_prof) {
    unsigned long long h[16];
    cudaSetDevice(s->device); cudaDeviceSynchronize();
    if (cudaMemcpy(h, s->uq_prof, sizeof(h), cudaMemcpyDeviceToHost) == cudaSuccess && h[15] > 0) {
      static const char* nm[9] = {"take tickets", "ring slots", "phase 0", "narrow phase", "rows (gather, constraint)", "controller", "actuation + acceleration", "solve", "integrate, obs, publish"};
      double tot = 0; for (int k = 0; k < 9; k++) tot += (double)h[k];
      fprintf(stderr, "[b2s] unit-queue stage profile over %llu block rounds (mean cycles per round, share):\n", h[15]);
      for (int k = 0; k < 9; k++) fprintf(stderr, "[b2s]   %-28s %9.0f  %5.1f %%\n", nm[k], (double)h[k] / (double)h[15], 100.0 * (double)h[k] / tot);
    }
  }
  if (s->slot >= 0 && g_slots[s->device & 63][s->slot] == s) g_slots[s->device & 63][s->slot] = nullptr;
  cudaSetDevice(s->device);
  cudaDeviceSynchronize();  // kernels of this handle may still be reading its buffers
  for (void* p : s->allocs) cudaFree(p);
  for (auto q : s->gstreams) cudaStreamDestroy(q);
  for (auto ev : s->gevents) cudaEventDestroy(ev);
  if (s->fork_event) cudaEventDestroy(s->fork_event);
  if (s->in_event) cudaEventDestroy(s->in_event);
  if (s->out_event) cudaEventDestroy(s->out_event);
  for (auto& kv : s->graphs) cudaGraphExecDestroy(kv.second);
  if (s->pstream) cudaStreamDestroy(s->pstream);
  delete s;
}

int b2s_set_stream(b2s_sim* s, void* stream) {
  if (!s) return fail(B2S_ERR_ARG, "null handle");
  s->stream = (cudaStream_t)stream;
  return B2S_OK;
}

int b2s_timeline(b2s_sim* s, int enable, double mean_us[8], int count[8]) {
  if (!s) return fail(B2S_ERR_ARG, "null handle");
  if (mean_us) for (int k = 0; k < 8; k++) mean_us[k] = s->tl_mean_us[k];
  if (count) for (int k = 0; k < 8; k++) count[k] = s->tl_count[k];
  if (enable >= 0) {
    s->timeline = enable ? (getenv("B2S_TIMELINE") ? 2 : 1) : 0;
    s->use_graph = enable ? 0 : (getenv("B2S_NO_GRAPH") ? 0 : 1);
  }
  return B2S_OK;
}

int b2s_set_profile(b2s_sim* s, int flag) { if (!s) return fail(B2S_ERR_ARG, "null handle"); s->profile = flag != 0; return B2S_OK; }

int b2s_set_export(b2s_sim* s, int flag) { if (!s) return fail(B2S_ERR_ARG, "null handle"); s->export_env_step = flag != 0; return B2S_OK; }

int64_t b2s_launch_count(const b2s_sim* s) { return s ? s->launches : 0; }

int b2s_array(b2s_sim* s, const char* name, void** dev_ptr, int* dtype, int* ndim, int64_t shape[4]) {
  if (!s || !name) return fail(B2S_ERR_ARG, "b2s_array: bad argument");
  auto it = s->arrays.find(name);
  if (it == s->arrays.end()) return fail(B2S_ERR_ARG, std::string("b2s_array: unknown array '") + name + "'");
  if (dev_ptr) *dev_ptr = it->second.ptr;
  if (dtype) *dtype = it->second.dtype;
  if (ndim) *ndim = it->second.ndim;
  if (shape) for (int k = 0; k < 4; k++) shape[k] = it->second.shape[k];
  return B2S_OK;
}

// upload this handle's descriptors into its constant-memory slot (only after a configuration change)
static int bind_constants(b2s_sim* s) {
  CUDA_TRY(cudaSetDevice(s->device));
  if (!s->dirty) return B2S_OK;
  const size_t k = (size_t)s->slot;
  if (s->precision == B2S_F32) {
    CUDA_TRY(cudaMemcpyToSymbolAsync(c_model_f, &s->mf, sizeof(s->mf), k * sizeof(s->mf), cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaMemcpyToSymbolAsync(c_state_f, &s->sf, sizeof(s->sf), k * sizeof(s->sf), cudaMemcpyHostToDevice, s->stream));
  } else {
    CUDA_TRY(cudaMemcpyToSymbolAsync(c_model_d, &s->md, sizeof(s->md), k * sizeof(s->md), cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaMemcpyToSymbolAsync(c_state_d, &s->sd, sizeof(s->sd), k * sizeof(s->sd), cudaMemcpyHostToDevice, s->stream));
  }
  CUDA_TRY(cudaMemcpyToSymbolAsync(c_lay, s->lay, sizeof(s->lay), k * sizeof(s->lay), cudaMemcpyHostToDevice, s->stream));
  CUDA_TRY(cudaMemcpyToSymbolAsync(c_cc, &s->ctrl, sizeof(s->ctrl), k * sizeof(s->ctrl), cudaMemcpyHostToDevice, s->stream));
  CUDA_TRY(cudaMemcpyToSymbolAsync(c_pio, s->pio, sizeof(s->pio), k * sizeof(s->pio), cudaMemcpyHostToDevice, s->stream));
  // the host structs must outlive the asynchronous copies only until they are enqueued: pageable-memory sources are staged
  s->dirty = 0;
  return B2S_OK;
}

// layouts depend on where the OSC controller runs; a change invalidates the captured graphs (they carry block shapes)
static int rebuild_layouts(b2s_sim* s) {
  const bool osc = s->ctrl.kind == B2S_CTRL_OSC_POSE || s->ctrl.kind == B2S_CTRL_OSC_POSITION;
  int want = (osc && (!s->ctrl_split || s->mode == 2)) ? 1 : 0;  // unit-queue mode runs the controller inside the unit
  if (want == s->osc_in_tail && s->smem0 != 0) return B2S_OK;
  s->uq_wpb = 0;
  s->osc_in_tail = want;
  try { build_layouts(s, s->ncg, s->hc_stride); } catch (const std::string& e) { return fail(B2S_ERR_MODEL, e); }
  int rc = choose_blocks(s);
  if (rc != B2S_OK) return rc;
  bool have_ws = s->precision == B2S_F32 ? s->sf.wsg != nullptr : s->sd.wsg != nullptr;
  if (have_ws) { cudaSetDevice(s->device); cudaDeviceSynchronize(); }
  for (auto& kv : s->graphs) cudaGraphExecDestroy(kv.second);
  s->graphs.clear();
  s->dirty = 1;
  return B2S_OK;
}

static int launch(b2s_sim* s, int phases, int nsub, const void* action = nullptr, const uint8_t* mask = nullptr) {
  int rc = bind_constants(s);
  if (rc != B2S_OK) return rc;
  int blocks = (s->n_env + s->wpb_fused - 1) / s->wpb_fused;
  if (s->precision == B2S_F32)
    step_kernel<float><<<blocks, s->wpb_fused * 32, s->smem_fused, s->stream>>>(phases, nsub, (const float*)action, s->slot, mask);
  else
    step_kernel<double><<<blocks, s->wpb_fused * 32, s->smem_fused, s->stream>>>(phases, nsub, (const double*)action, s->slot, mask);
  s->launches++;
  CUDA_TRY(cudaGetLastError());
  return B2S_OK;
}

}  // extern "C"

// enqueue the launches of `nsub` substeps for every environment group; `q0` is the stream the caller forks from / joins to
// the launches of `nsub` substeps of ONE environment group on stream q
template <typename R> static int enqueue_group(b2s_sim* s, DState<R>& st, int phases, int nsub, const R* action, int gi, int G, cudaStream_t q) {
  const int epaw = (EPA_PIPE_WORDS + (s->precision == B2S_F32 ? s->mf.stage_cap : s->md.stage_cap)) * (int)sizeof(R);
  const int p1smem = std::max(epaw, (int)osc_smem_bytes<R>());  // one block shape for the three roles of phase 1
  const bool tiered = s->mc_small < s->maxcon || s->me_small < s->maxefc;
  {
    int e0 = (int)((long long)s->n_env * gi / G), e1 = (int)((long long)s->n_env * (gi + 1) / G);
    Grp g{e0, e1 - e0, gi, 0, s->slot};
    int blocks0 = (g.nenv + s->wpb0 - 1) / s->wpb0, blocks5 = (g.nenv + s->wpb5s - 1) / s->wpb5s;
    int blocksL = std::min((g.nenv + s->wpb5l - 1) / s->wpb5l, 2 * 148);  // large tier: warps claim overflowed environments
    int nA = g.nenv * st.cl_maxa, nG = g.nenv * st.cl_maxg;
    // convex role: one warp per block, items claimed through a counter.  ~1.6 items per environment are queued per substep (Lift), most of
    // them dismissed in a few microseconds: half a block per environment keeps every slow item on its own warp without flooding the
    // block scheduler with thousands of empty blocks per launch (B2S_CVX_BLOCKS overrides)
    int cvx_blocks = std::max(148, g.nenv / 2);
    if (const char* v = getenv("B2S_CVX_BLOCKS")) { int x = atoi(v); if (x > 0) cvx_blocks = x; }
    const bool ctrl_ext = (phases & PH_CTRL_EXT) != 0;
    // B2S_TIMELINE=1 (with B2S_NO_GRAPH=1): timing events between the launches, per-kernel means on stderr (debug aid)
    auto mark = [&](int type) {
      if (!s->timeline) return;
      cudaEvent_t ev; cudaEventCreate(&ev); cudaEventRecord(ev, q);
      s->tl_events.push_back({gi, type, ev});
    };
    for (int sub = 0; sub < nsub; sub++) {
      g.sub = sub;
      mark(0);
      CUDA_TRY(cudaMemsetAsync(st.cl_cnt + 8 * gi, 0, 8 * sizeof(int), q));
      mark(1);
      phase0_kernel<R><<<blocks0, s->wpb0 * 32, s->smem0, q>>>(phases, g);
      mark(2);
      // phase 1: convex narrow phase | controller | analytic narrow phase as block roles of ONE launch (no forks in the graph).
      // Upper bounds of the candidate counts size the grid; warps / threads beyond the device-side counts exit at once.
      {
        P1Cfg c{(s->debug_skip & 2) ? 0 : std::min(nG, cvx_blocks), ctrl_ext ? (g.nenv + OSC_TPB - 1) / OSC_TPB : 0, sub};
        int nAb = (s->debug_skip & 1) ? 0 : (nA + 31) / 32;
        if (c.nG + c.nC + nAb > 0) phase1_kernel<R><<<c.nG + c.nC + nAb, 32, p1smem, q>>>(action, g, c);
      }
      mark(4);
      tail_kernel<R><<<blocks5, s->wpb5s * 32, s->smem5s, q>>>(phases, nsub, action, g, 0);
      mark(5);
      if (tiered) {
        tail_kernel<R><<<blocksL, s->wpb5l * 32, s->smem5l, q>>>(phases, nsub, action, g, 1);
        mark(6);
      }
    }
  }
  CUDA_TRY(cudaGetLastError());
  return B2S_OK;
}

// all groups, forked from / joined to q0 (eager mode and the single-graph capture)
template <typename R> static int enqueue_pipeline(b2s_sim* s, DState<R>& st, int phases, int nsub, const R* action, cudaStream_t q0) {
  int G = s->ngroups;
  if (G > s->n_env) G = s->n_env;
  CUDA_TRY(cudaEventRecord(s->fork_event, q0));
  for (int gi = 0; gi < G; gi++) {
    cudaStream_t q = G == 1 ? q0 : s->gstreams[gi];
    if (G > 1) CUDA_TRY(cudaStreamWaitEvent(q, s->fork_event, 0));
    int rc = enqueue_group<R>(s, st, phases, nsub, action, gi, G, q);
    if (rc != B2S_OK) return rc;
    if (G > 1) {
      CUDA_TRY(cudaEventRecord(s->gevents[gi], q));
      CUDA_TRY(cudaStreamWaitEvent(q0, s->gevents[gi], 0));
    }
  }
  return B2S_OK;
}

// global workspace rows + per-environment candidate tables / narrow-phase output slots (pipeline and unit-queue modes)
template <typename R> static int ensure_ws(b2s_sim* s, DState<R>& st) {
  if (!st.wsg) {
    R* p = nullptr;
    if (cudaMalloc(&p, (size_t)s->n_env * s->lay[LAY_ROW].total * sizeof(R)) != cudaSuccess) return fail(B2S_ERR_CUDA, "cudaMalloc(pipeline workspace) failed");
    cudaMemsetAsync(p, 0, (size_t)s->n_env * s->lay[LAY_ROW].total * sizeof(R), s->stream);
    s->allocs.push_back(p);
    st.wsg = p;
    size_t ne = (size_t)s->n_env;
    st.cl_cnt = dev_zeros<int>(s, 8 * 64);
    st.ovf_list = dev_zeros<int>(s, ne);
    // candidate capacity per environment: small models keep small grids (the narrow-phase grids are sized by these bounds)
    st.cl_maxa = s->maxcon <= 32 ? 8 : (s->maxcon <= 48 ? 16 : CL_MAXA);
    st.cl_maxg = s->maxcon <= 32 ? 16 : CL_MAXG;
    if (const char* v = getenv("B2S_CL_MAXA")) { int x = atoi(v); if (x >= 1 && x <= CL_MAXA) st.cl_maxa = x; }
    if (const char* v = getenv("B2S_CL_MAXG")) { int x = atoi(v); if (x >= 1 && x <= CL_MAXG) st.cl_maxg = x; }
    st.cl_listA = dev_zeros<int>(s, ne * st.cl_maxa); st.cl_listG = dev_zeros<int>(s, ne * st.cl_maxg);
    st.cl_outA = dev_zeros<R>(s, ne * st.cl_maxa * CL_RECA); st.cl_outG = dev_zeros<R>(s, ne * st.cl_maxg * 8);
    st.cl_env = dev_zeros<int>(s, ne * CL_ENVW(st));
    st.gjk_cache = getenv("B2S_NO_GJK_CACHE") ? nullptr : dev_zeros<R>(s, ne * (size_t)(s->precision == B2S_F32 ? s->mf.npair : s->md.npair) * 3);
    s->action_buf = dev_zeros<R>(s, ne * 16);
#ifdef B2S_INSTR
    st.st_begin = dev_zeros<unsigned long long>(s, 64 * 32 * 8); st.st_end = dev_zeros<unsigned long long>(s, 64 * 32 * 8);
    st.stats = dev_zeros<int>(s, 512); st.cyc = dev_zeros<float>(s, ne * 64);
    s->arrays["st_begin"] = ArrayInfo{st.st_begin, B2S_I64, 1, {64 * 32 * 8, 0, 0, 0}};
    s->arrays["st_end"] = ArrayInfo{st.st_end, B2S_I64, 1, {64 * 32 * 8, 0, 0, 0}};
    s->arrays["stats"] = ArrayInfo{st.stats, B2S_I32, 1, {512, 0, 0, 0}};
    s->arrays["cyc"] = ArrayInfo{st.cyc, B2S_F32, 3, {(int64_t)ne, 32, 2, 0}};
    st.slowlog = dev_zeros<int>(s, 64 * 12);
    s->arrays["slowlog"] = ArrayInfo{st.slowlog, B2S_I32, 2, {64, 12, 0, 0}};
#endif
    s->dirty = 1;
  }
  return B2S_OK;
}

template <typename R> static int launch_pipeline_t(b2s_sim* s, DState<R>& st, int phases, int nsub, const R* action) {
  { int rc0 = ensure_ws<R>(s, st); if (rc0 != B2S_OK) return rc0; }
#ifdef B2S_INSTR
  cudaMemsetAsync(st.st_begin, 0xff, sizeof(unsigned long long) * 64 * 32 * 8, s->stream);
  cudaMemsetAsync(st.st_end, 0, sizeof(unsigned long long) * 64 * 32 * 8, s->stream);
#endif
  phases |= PH_WORKLIST;
  const bool osc = s->ctrl.kind == B2S_CTRL_OSC_POSE || s->ctrl.kind == B2S_CTRL_OSC_POSITION;
  if ((phases & PH_CTRL) && osc && s->ctrl_split) phases |= PH_CTRL_EXT;
  { int rc2 = rebuild_layouts(s); if (rc2 != B2S_OK) return rc2; }  // before the descriptors are (re)uploaded
  int rc = bind_constants(s);
  if (rc != B2S_OK) return rc;
  int G = s->ngroups;
  if (G > s->n_env) G = s->n_env;
  while ((int)s->gstreams.size() < G) {
    cudaStream_t st2; cudaEvent_t ev;
    CUDA_TRY(cudaStreamCreateWithFlags(&st2, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    s->gstreams.push_back(st2); s->gevents.push_back(ev);
  }
  if (!s->fork_event) {
    CUDA_TRY(cudaEventCreateWithFlags(&s->fork_event, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&s->in_event, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&s->out_event, cudaEventDisableTiming));
    CUDA_TRY(cudaStreamCreateWithFlags(&s->pstream, cudaStreamNonBlocking));
  }
  // kernels per group-substep: phase 0, analytic + convex narrow phase, then the merged tail (or phases 2, [3], 4)
  const bool tiered = s->mc_small < s->maxcon || s->me_small < s->maxefc;
  int launches_per_call = G * nsub * (3 + (tiered ? 1 : 0));  // phase 0, phase 1 (narrow phase + controller), tail, [tail large tier]
  if (!s->use_graph) {
    rc = enqueue_pipeline<R>(s, st, phases, nsub, action, s->stream);
    s->launches += launches_per_call;
    if (s->timeline && rc == B2S_OK) {
      cudaStreamSynchronize(s->stream);
      static const char* names[8] = {"(prev->memset)", "memset", "P0", "narrowA", "narrowG", "tail", "tail(large tier)", "-"};
      double sum[8] = {0}; int cnt[8] = {0};
      for (size_t i = 1; i < s->tl_events.size(); i++) {
        auto &a = s->tl_events[i - 1], &b = s->tl_events[i];
        if (a.group != b.group) continue;
        float ms = 0; cudaEventElapsedTime(&ms, a.ev, b.ev);
        sum[b.type] += ms; cnt[b.type]++;
      }
      double tot = 0;
      for (int k = 0; k < 8; k++) {
        s->tl_mean_us[k] = cnt[k] ? 1e3 * sum[k] / cnt[k] : 0.0; s->tl_count[k] = cnt[k];
        if (cnt[k]) tot += sum[k];
        if (cnt[k] && s->timeline > 1) fprintf(stderr, "[timeline] %-16s n=%4d mean %8.1f us\n", names[k], cnt[k], s->tl_mean_us[k]);
      }
      if (s->timeline > 1) fprintf(stderr, "[timeline] sum over one call %.3f ms (all groups)\n", tot);
      for (auto& t : s->tl_events) cudaEventDestroy(t.ev);
      s->tl_events.clear();
    }
    return rc;
  }
  // CUDA-graph replay: the launch sequence of one call (G groups x nsub substeps x 6 kernels) is captured once per
  // (phases, nsub) on an internal stream; the action rows are staged into a fixed buffer so kernel arguments never change
  const R* act_in = action;
  if (action) {
    int ad = s->ctrl.action_dim > 0 ? s->ctrl.action_dim : 1;
    if (ad > 16) return fail(B2S_ERR_UNSUPPORTED, "action_dim > 16");
    CUDA_TRY(cudaMemcpyAsync(s->action_buf, action, (size_t)s->n_env * ad * sizeof(R), cudaMemcpyDeviceToDevice, s->stream));
    act_in = (const R*)s->action_buf;
  }
  CUDA_TRY(cudaEventRecord(s->in_event, s->stream));
  if (s->graph_per_group && G > 1) {
    // one graph per environment group, each a plain chain replayed on its own stream: the groups' chains then overlap freely
    // (inside ONE graph, parallel branches were observed to share a limited number of execution lanes)
    for (int gi = 0; gi < G; gi++) {
      long long key = ((long long)phases << 28) | ((long long)(gi + 1) << 20) | (long long)nsub << 4 | (action ? 1 : 0);
      cudaStream_t q = s->gstreams[gi];
      auto it = s->graphs.find(key);
      if (it == s->graphs.end()) {
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        CUDA_TRY(cudaStreamBeginCapture(q, cudaStreamCaptureModeRelaxed));
        rc = enqueue_group<R>(s, st, phases, nsub, act_in, gi, G, q);
        cudaError_t ce = cudaStreamEndCapture(q, &graph);
        if (rc != B2S_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (ce != cudaSuccess) return fail(B2S_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
        CUDA_TRY(cudaGraphInstantiate(&exec, graph, 0));
        cudaGraphDestroy(graph);
        it = s->graphs.emplace(key, exec).first;
      }
      CUDA_TRY(cudaStreamWaitEvent(q, s->in_event, 0));
      CUDA_TRY(cudaGraphLaunch(it->second, q));
      CUDA_TRY(cudaEventRecord(s->gevents[gi], q));
      CUDA_TRY(cudaStreamWaitEvent(s->stream, s->gevents[gi], 0));
    }
    s->launches += launches_per_call;
    return B2S_OK;
  }
  CUDA_TRY(cudaStreamWaitEvent(s->pstream, s->in_event, 0));
  long long key = ((long long)phases << 28) | (long long)nsub << 4 | (action ? 1 : 0);
  auto it = s->graphs.find(key);
  if (it == s->graphs.end()) {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    CUDA_TRY(cudaStreamBeginCapture(s->pstream, cudaStreamCaptureModeRelaxed));
    rc = enqueue_pipeline<R>(s, st, phases, nsub, act_in, s->pstream);
    cudaError_t ce = cudaStreamEndCapture(s->pstream, &graph);
    if (rc != B2S_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess) return fail(B2S_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
    CUDA_TRY(cudaGraphInstantiate(&exec, graph, 0));
    cudaGraphDestroy(graph);
    it = s->graphs.emplace(key, exec).first;
  }
  CUDA_TRY(cudaGraphLaunch(it->second, s->pstream));
  CUDA_TRY(cudaEventRecord(s->out_event, s->pstream));
  CUDA_TRY(cudaStreamWaitEvent(s->stream, s->out_event, 0));
  s->launches += launches_per_call;
  return B2S_OK;
}
static int launch_pipeline(b2s_sim* s, int phases, int nsub, const void* action) {
  return s->precision == B2S_F32 ? launch_pipeline_t<float>(s, s->sf, phases, nsub, (const float*)action)
                                 : launch_pipeline_t<double>(s, s->sd, phases, nsub, (const double*)action);
}

// ---- unit-queue mode: one persistent kernel per control step (b2s_unit.cuh)
template <typename R> static int launch_unit_t(b2s_sim* s, DState<R>& st, int phases, int nsub, const R* action) {
  { int rc0 = ensure_ws<R>(s, st); if (rc0 != B2S_OK) return rc0; }
  { int rc2 = rebuild_layouts(s); if (rc2 != B2S_OK) return rc2; }
  const int total = s->n_env * nsub;
  if ((long long)s->n_env * nsub > (1ll << 30)) return fail(B2S_ERR_UNSUPPORTED, "unit-queue mode: n_env * nsub too large");
  if (total > s->uq_cap) {
    cudaStreamSynchronize(s->stream);
    int* p = nullptr;
    if (cudaMalloc(&p, sizeof(int) * (2 * (size_t)total + 8)) != cudaSuccess) return fail(B2S_ERR_CUDA, "cudaMalloc(unit ring) failed");
    s->allocs.push_back(p);
    s->uq_ring = p; s->uq_ovf = p + total; s->uq_ctr = p + 2 * (size_t)total; s->uq_cap = total;
  }
  if (s->uq_wpb == 0) {
    // block shape: the warp's one workspace area holds phase 0's layout, then the EPA polytope + vertex staging, then the small tail tier
    const size_t rsz = sizeof(R);
    const bool tiered = s->mc_small < s->maxcon || s->me_small < s->maxefc;
    int stride = std::max(std::max(s->lay[LAY_P0].total, s->lay[LAY_TS].total), EPA_PIPE_WORDS + 24 + 384);
    stride = (stride + 3) & ~3;
    int stride_l = (s->lay[LAY_TL].total + 3) & ~3;
    CUDA_TRY(optin_max_smem(unit_kernel<R>, s->device));
    int best_w = 0, best_b = 0, best = 0;
    int wcap = B2S_LBU_THREADS / 32;
    if (const char* v = getenv("B2S_UNIT_WPB")) { int x = atoi(v); if (x >= 1 && x <= wcap) wcap = x; }
    for (int w = wcap; w >= 1; w--) {
      size_t sm = std::max((size_t)w * stride, tiered ? (size_t)stride_l : 0) * rsz;
      if (sm > 226 * 1024) continue;
      int b = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, unit_kernel<R>, w * 32, sm) != cudaSuccess) { cudaGetLastError(); continue; }
      if (w * b > best) { best = w * b; best_w = w; best_b = b; }
    }
    if (best == 0) return fail(B2S_ERR_UNSUPPORTED, "unit-queue mode: workspace does not fit shared memory");
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, s->device));
    s->uq_wpb = best_w; s->uq_bps = best_b; s->uq_stride = stride; s->uq_stride_large = stride_l;
    s->uq_smem = std::max((size_t)best_w * stride, tiered ? (size_t)stride_l : 0) * rsz;
    s->uq_wpb_large = tiered ? std::min(best_w, (int)(s->uq_smem / rsz / stride_l)) : 0;
    int slots = best_b * prop.multiProcessorCount;
    int nl = 0;
    if (tiered) {
      nl = std::max(1, std::min(slots / 64, 12));  // a large-role block occupies a block slot (a whole SM at one block per SM)
      if (const char* v = getenv("B2S_UNIT_LARGE_BLOCKS")) { int x = atoi(v); if (x >= 1 && x < slots) nl = x; }
    }
    s->uq_nlarge = nl;
    if (best_w > s->n_env) best_w = s->n_env;  // the lockstep rounds need n_env >= warps per block (b2s_unit.cuh)
    s->uq_wpb = best_w;
    int small_blocks = std::min(std::max(slots - nl, 1), (s->n_env + best_w - 1) / best_w);
    if (const char* v = getenv("B2S_UNIT_BLOCKS")) { int x = atoi(v); if (x >= 1) small_blocks = x; }
    s->uq_grid = nl + small_blocks;
    if (getenv("B2S_VERBOSE"))
      fprintf(stderr, "[b2s] unit-queue: %d warps/block x %d blocks/SM, %d words/warp (large role: %d words, %d warps/block, %d blocks), grid %d, smem %zu B\n",
              best_w, best_b, stride, stride_l, s->uq_wpb_large, nl, s->uq_grid, s->uq_smem);
  }
  int rc = bind_constants(s);
  if (rc != B2S_OK) return rc;
  if (getenv("B2S_UNIT_PROF") && !s->uq_prof) s->uq_prof = dev_zeros<unsigned long long>(s, 16);
  int ubar = 4;
  if (const char* v = getenv("B2S_UNIT_BARRIERS")) ubar = atoi(v);
  UnitQ q{s->uq_ring, s->uq_ovf, s->uq_ctr, total, s->uq_nlarge, s->uq_wpb_large, s->uq_stride, s->uq_stride_large, s->uq_prof, ubar};
  unit_init_kernel<R><<<(total + 255) / 256, 256, 0, s->stream>>>(q, s->n_env);
  unit_kernel<R><<<s->uq_grid, s->uq_wpb * 32, s->uq_smem, s->stream>>>(phases, nsub, action, s->slot, q);
  unit_check_kernel<R><<<8, 256, 0, s->stream>>>(q, s->slot);
  s->launches += 3;
  CUDA_TRY(cudaGetLastError());
  if (getenv("B2S_UNIT_DEBUG")) {
    int c[8];
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    CUDA_TRY(cudaMemcpy(c, s->uq_ctr, sizeof(c), cudaMemcpyDeviceToHost));
    fprintf(stderr, "[b2s] unit ctr: head %d tail %d done %d ovf_head %d ovf_tail %d | watchdog ticket %d tail_then %d flag %d (total %d)\n", c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7], total);
    if (c[7]) return fail(B2S_ERR_CUDA, "unit-queue watchdog: a ticket was never produced");
  }
  return B2S_OK;
}
static int launch_unit(b2s_sim* s, int phases, int nsub, const void* action) {
  return s->precision == B2S_F32 ? launch_unit_t<float>(s, s->sf, phases, nsub, (const float*)action)
                                 : launch_unit_t<double>(s, s->sd, phases, nsub, (const double*)action);
}

extern "C" {

/* Host-only: the workspace layouts libb2s would build for a model of the given dimensions (no device needed).  out_words[5] receives
 * words per warp / row of LAY_FULL (incl. the EPA area), LAY_P0, LAY_TS, LAY_TL, LAY_ROW; out_layouts (may be NULL) receives the five
 * WSLayout structs as ints, out_pio (may be NULL) B2S_NPIO x (nload, nstore, load_words, then 12 + 12 regions x 4 ints). */
int b2s_debug_layouts(int nq, int nv, int nu, int nbody, int ncg, int nsite, int hc_stride, int maxcon, int maxefc, int mc_small,
                      int me_small, int osc_in_tail, int* out_words, int* out_layouts, int* out_pio) {
  b2s_sim tmp;
  tmp.nq = nq; tmp.nv = nv; tmp.nu = nu; tmp.nbody = nbody; tmp.nsite = nsite; tmp.maxcon = maxcon; tmp.maxefc = maxefc;
  tmp.mc_small = mc_small; tmp.me_small = me_small; tmp.osc_in_tail = osc_in_tail;
  try { build_layouts(&tmp, ncg, hc_stride); } catch (const std::string& e) { return fail(B2S_ERR_MODEL, e); }
  if (out_words) {
    out_words[0] = tmp.lay[LAY_FULL].fused_stride;
    for (int k = 1; k < B2S_NLAY; k++) out_words[k] = tmp.lay[k].total;
  }
  if (out_layouts) memcpy(out_layouts, tmp.lay, sizeof(tmp.lay));
  if (out_pio) {
    for (int k = 0; k < B2S_NPIO; k++) {
      int* o = out_pio + k * (3 + 2 * B2S_MAXREG * 4);
      o[0] = tmp.pio[k].nload; o[1] = tmp.pio[k].nstore; o[2] = tmp.pio[k].load_words;
      memcpy(o + 3, tmp.pio[k].load, sizeof(Region) * B2S_MAXREG);
      memcpy(o + 3 + 4 * B2S_MAXREG, tmp.pio[k].store, sizeof(Region) * B2S_MAXREG);
    }
  }
  return B2S_OK;
}

int b2s_set_mode(b2s_sim* s, int mode) {
  if (!s || mode < 0 || mode > 2) return fail(B2S_ERR_ARG, "b2s_set_mode: mode must be 0 (fused), 1 (pipeline) or 2 (unit queue)");
  s->mode = mode;
  const char* eg = getenv("B2S_GROUPS");
  if (eg) { int v = atoi(eg); if (v >= 1 && v <= 64) s->ngroups = v; }
  if (getenv("B2S_NO_GRAPH")) s->use_graph = 0;
  if (const char* ds = getenv("B2S_DEBUG_SKIP")) s->debug_skip = atoi(ds);
  if (getenv("B2S_TIMELINE")) { s->timeline = 2; s->use_graph = 0; }
  if (const char* v = getenv("B2S_CTRL_SPLIT")) s->ctrl_split = atoi(v) != 0;
  if (const char* v = getenv("B2S_GRAPH_PER_GROUP")) s->graph_per_group = atoi(v) != 0;
  return B2S_OK;
}

static void clear_warm_start(b2s_sim* s, const uint8_t* mask) {
  int npair = s->precision == B2S_F32 ? s->mf.npair : s->md.npair;
  bool have = s->precision == B2S_F32 ? s->sf.gjk_cache != nullptr : s->sd.gjk_cache != nullptr;
  if (!have || npair == 0) return;
  size_t total = (size_t)s->n_env * npair * 3;
  int blocks = (int)((total + 255) / 256);
  if (s->precision == B2S_F32) cache_reset_kernel<float><<<blocks, 256, 0, s->stream>>>(mask, s->slot);
  else cache_reset_kernel<double><<<blocks, 256, 0, s->stream>>>(mask, s->slot);
  s->launches++;
}

int b2s_reset(b2s_sim* s, const uint8_t* mask) {
  if (!s) return fail(B2S_ERR_ARG, "null handle");
  { int rc = bind_constants(s); if (rc != B2S_OK) return rc; }
  if (s->precision == B2S_F32) reset_kernel<float><<<(s->n_env + 127) / 128, 128, 0, s->stream>>>(mask, s->slot);
  else reset_kernel<double><<<(s->n_env + 127) / 128, 128, 0, s->stream>>>(mask, s->slot);
  clear_warm_start(s, mask);
  s->launches++;
  CUDA_TRY(cudaGetLastError());
  return B2S_OK;
}

int b2s_forward(b2s_sim* s) { return s ? launch(s, PH_STEP1 | PH_STEP2 | PH_NOINTEGRATE | PH_EXPORT | (s->has_obs ? PH_OBS : 0), 1) : fail(B2S_ERR_ARG, "null handle"); }
int b2s_step1(b2s_sim* s) { return s ? launch(s, PH_STEP1 | PH_EXPORT, 1) : fail(B2S_ERR_ARG, "null handle"); }
int b2s_step2(b2s_sim* s) { return s ? launch(s, PH_STEP1 | PH_STEP2 | PH_EXPORT, 1) : fail(B2S_ERR_ARG, "null handle"); }
int b2s_step(b2s_sim* s, int n) {
  if (!s || n < 1) return fail(B2S_ERR_ARG, "b2s_step: bad argument");
  if (s->mode == 2) return launch_unit(s, PH_STEP1 | PH_STEP2, n, nullptr);
  if (s->mode == 1) return launch_pipeline(s, PH_STEP1 | PH_STEP2, n, nullptr);
  return launch(s, PH_STEP1 | PH_STEP2, n);
}

int b2s_jac_site(b2s_sim* s, int site_id, void* jacp, void* jacr) {
  if (!s || site_id < 0 || site_id >= s->nsite) return fail(B2S_ERR_ARG, "b2s_jac_site: bad argument");
  { int rc = bind_constants(s); if (rc != B2S_OK) return rc; }
  int threads = 128, blocks = (s->n_env * s->nv + threads - 1) / threads;
  if (s->precision == B2S_F32) jac_site_kernel<float><<<blocks, threads, 0, s->stream>>>(site_id, (float*)jacp, (float*)jacr, s->slot);
  else jac_site_kernel<double><<<blocks, threads, 0, s->stream>>>(site_id, (double*)jacp, (double*)jacr, s->slot);
  s->launches++;
  CUDA_TRY(cudaGetLastError());
  return B2S_OK;
}

int b2s_get_state(b2s_sim* s, void* out) {
  if (!s || !out) return fail(B2S_ERR_ARG, "b2s_get_state: bad argument");
  { int rc = bind_constants(s); if (rc != B2S_OK) return rc; }
  size_t total = (size_t)s->n_env * (1 + s->nq + s->nv);
  int blocks = (int)((total + 255) / 256);
  if (s->precision == B2S_F32) state_io_kernel<float><<<blocks, 256, 0, s->stream>>>((float*)out, 0, s->slot);
  else state_io_kernel<double><<<blocks, 256, 0, s->stream>>>((double*)out, 0, s->slot);
  s->launches++;
  CUDA_TRY(cudaGetLastError());
  return B2S_OK;
}
int b2s_set_state(b2s_sim* s, const void* in) {
  if (!s || !in) return fail(B2S_ERR_ARG, "b2s_set_state: bad argument");
  { int rc = bind_constants(s); if (rc != B2S_OK) return rc; }
  size_t total = (size_t)s->n_env * (1 + s->nq + s->nv);
  int blocks = (int)((total + 255) / 256);
  if (s->precision == B2S_F32) state_io_kernel<float><<<blocks, 256, 0, s->stream>>>((float*)in, 1, s->slot);
  else state_io_kernel<double><<<blocks, 256, 0, s->stream>>>((double*)in, 1, s->slot);
  s->launches++;
  CUDA_TRY(cudaGetLastError());
  return B2S_OK;
}
int b2s_name2id(const b2s_sim* s, const char* type, const char* name) {
  if (!s || !type || !name) return -1;
  auto it = s->names.find(type);
  if (it == s->names.end()) return -1;
  for (size_t i = 0; i < it->second.size(); i++) if (it->second[i] == name) return (int)i;
  return -1;
}
const char* b2s_id2name(const b2s_sim* s, const char* type, int id) {
  if (!s || !type) return nullptr;
  auto it = s->names.find(type);
  if (it == s->names.end() || id < 0 || id >= (int)it->second.size()) return nullptr;
  return it->second[id].c_str();
}
int b2s_full_m(b2s_sim* s, void* out) {
  if (!s || !out) return fail(B2S_ERR_ARG, "b2s_full_m: bad argument");
  size_t rsz = s->precision == B2S_F32 ? 4 : 8;
  const void* src = s->precision == B2S_F32 ? (const void*)s->sf.qM : (const void*)s->sd.qM;
  CUDA_TRY(cudaMemcpyAsync(out, src, (size_t)s->n_env * s->nv * s->nv * rsz, cudaMemcpyDeviceToDevice, s->stream));
  return B2S_OK;
}
static int jac_point(b2s_sim* s, int kind, int id, void* jacp, void* jacr) {
  { int rc = bind_constants(s); if (rc != B2S_OK) return rc; }
  int threads = 128, blocks = (s->n_env * s->nv + threads - 1) / threads;
  if (s->precision == B2S_F32) jac_point_kernel<float><<<blocks, threads, 0, s->stream>>>(kind, id, (float*)jacp, (float*)jacr, s->slot);
  else jac_point_kernel<double><<<blocks, threads, 0, s->stream>>>(kind, id, (double*)jacp, (double*)jacr, s->slot);
  s->launches++;
  CUDA_TRY(cudaGetLastError());
  return B2S_OK;
}
int b2s_jac_body(b2s_sim* s, int body_id, void* jacp, void* jacr) {
  if (!s || body_id < 0 || body_id >= s->nbody) return fail(B2S_ERR_ARG, "b2s_jac_body: bad argument");
  return jac_point(s, 0, body_id, jacp, jacr);
}
int b2s_jac_geom(b2s_sim* s, int geom_id, void* jacp, void* jacr) {
  if (!s || geom_id < 0 || geom_id >= s->ngeom) return fail(B2S_ERR_ARG, "b2s_jac_geom: bad argument");
  return jac_point(s, 1, geom_id, jacp, jacr);
}

int b2s_ctrl_config(b2s_sim* s, const b2s_ctrl_cfg* c) {
  if (!s || !c) return fail(B2S_ERR_ARG, "b2s_ctrl_config: bad argument");
  if (c->kind != B2S_CTRL_OSC_POSE && c->kind != B2S_CTRL_JOINT_VELOCITY && c->kind != B2S_CTRL_JOINT_POSITION &&
      c->kind != B2S_CTRL_JOINT_TORQUE && c->kind != B2S_CTRL_OSC_POSITION && c->kind != B2S_CTRL_NONE)
    return fail(B2S_ERR_UNSUPPORTED, "controller kind not implemented");
  if (c->n_arm > 8 || c->n_grip > 4) return fail(B2S_ERR_ARG, "b2s_ctrl_config: too many joints");
  CtrlCfgDev& d = s->ctrl;
  d.kind = c->kind; d.action_dim = c->action_dim; d.n_arm = c->n_arm; d.eef_site = c->eef_site; d.base_site = c->base_site;
  d.n_grip = c->n_grip; d.uncouple = c->uncouple_pos_ori;
  for (int i = 0; i < 8; i++) { d.arm_dof[i] = c->arm_dof[i]; d.arm_qpos[i] = c->arm_qpos[i]; d.arm_act[i] = c->arm_act[i]; }
  for (int i = 0; i < 4; i++) { d.grip_act[i] = c->grip_act[i]; d.grip_sign[i] = c->grip_sign[i]; }
  d.grip_speed = c->grip_speed; d.null_kp = c->null_kp;
  for (int i = 0; i < 6; i++) {
    d.kp[i] = c->kp[i]; d.kd[i] = 2.0 * sqrt(c->kp[i]) * c->damping_ratio[i];
    d.input_max[i] = c->input_max[i]; d.input_min[i] = c->input_min[i];
    d.output_max[i] = c->output_max[i]; d.output_min[i] = c->output_min[i];
  }
  for (int i = 0; i < 8; i++) {
    d.jv_kp[i] = c->jv_kp[i]; d.jv_ki[i] = c->jv_ki[i]; d.jv_kd[i] = c->jv_kd[i];
    d.jv_in_max[i] = c->jv_in_max[i]; d.jv_in_min[i] = c->jv_in_min[i]; d.jv_out_max[i] = c->jv_out_max[i]; d.jv_out_min[i] = c->jv_out_min[i];
  }
  d.jv_vel_lo = c->jv_vel_lo; d.jv_vel_hi = c->jv_vel_hi; d.jv_use_vel_limits = c->jv_use_vel_limits; d.jv_torque_comp = c->jv_torque_comp;
  s->has_ctrl = c->kind != B2S_CTRL_NONE;
  s->dirty = 1;
  return B2S_OK;
}

int b2s_ctrl_reset(b2s_sim* s, const uint8_t* mask) {
  if (!s || !s->has_ctrl) return fail(B2S_ERR_ARG, "b2s_ctrl_reset: controller not configured");
  { int rc = bind_constants(s); if (rc != B2S_OK) return rc; }
  int threads = 128, blocks = (s->n_env + threads - 1) / threads;
  if (s->precision == B2S_F32) ctrl_reset_kernel<float><<<blocks, threads, 0, s->stream>>>(mask, s->slot);
  else ctrl_reset_kernel<double><<<blocks, threads, 0, s->stream>>>(mask, s->slot);
  clear_warm_start(s, mask);  // an environment whose controller is rebuilt starts a new episode
  s->launches++;
  CUDA_TRY(cudaGetLastError());
  return B2S_OK;
}

int b2s_reset_envs(b2s_sim* s, const uint8_t* mask, const void* qpos_new) {
  if (!s) return fail(B2S_ERR_ARG, "null handle");
  { int rc = bind_constants(s); if (rc != B2S_OK) return rc; }
  int threads = 128, blocks = (s->n_env + threads - 1) / threads;
  if (s->precision == B2S_F32) reset_envs_kernel<float><<<blocks, threads, 0, s->stream>>>(mask, (const float*)qpos_new, s->slot);
  else reset_envs_kernel<double><<<blocks, threads, 0, s->stream>>>(mask, (const double*)qpos_new, s->slot);
  s->launches++;
  CUDA_TRY(cudaGetLastError());
  int rc = launch(s, PH_STEP1 | PH_STEP2 | PH_NOINTEGRATE | PH_EXPORT | (s->has_obs ? PH_OBS : 0), 1, nullptr, mask);
  if (rc != B2S_OK) return rc;
  if (s->has_ctrl) return b2s_ctrl_reset(s, mask);
  clear_warm_start(s, mask);
  return B2S_OK;
}

}  // extern "C"
template <typename R> static int body_pose_override_t(b2s_sim* s, DState<R>& st, int body) {
  for (int k = 0; k < st.n_ov; k++) if (st.ov_body[k] == body) return B2S_OK;
  if (st.n_ov >= 4) return fail(B2S_ERR_UNSUPPORTED, "b2s_body_pose_override: at most 4 bodies per handle");
  std::vector<R> hp((size_t)s->n_env * 3), hq((size_t)s->n_env * 4);
  for (int e = 0; e < s->n_env; e++) {
    for (int k = 0; k < 3; k++) hp[(size_t)e * 3 + k] = (R)s->xpos0_h[3 * body + k];
    for (int k = 0; k < 4; k++) hq[(size_t)e * 4 + k] = (R)s->xquat0_h[4 * body + k];
  }
  const int k = st.n_ov;
  try { st.ov_pos[k] = dev_upload(s, hp); st.ov_quat[k] = dev_upload(s, hq); } catch (const std::string& e) { return fail(B2S_ERR_CUDA, e); }
  st.ov_body[k] = body;
  st.n_ov = k + 1;
  const int code = s->precision == B2S_F32 ? B2S_F32 : B2S_F64;
  s->arrays["body_xpos_ov:" + std::to_string(body)] = ArrayInfo{st.ov_pos[k], code, 2, {s->n_env, 3, 0, 0}};
  s->arrays["body_xquat_ov:" + std::to_string(body)] = ArrayInfo{st.ov_quat[k], code, 2, {s->n_env, 4, 0, 0}};
  s->dirty = 1;
  return B2S_OK;
}
extern "C" {
int b2s_body_pose_override(b2s_sim* s, int body_id) {
  if (!s || body_id <= 0 || body_id >= s->nbody) return fail(B2S_ERR_ARG, "b2s_body_pose_override: bad argument");
  if (s->body_weldid_h[body_id] != 0) return fail(B2S_ERR_UNSUPPORTED, "b2s_body_pose_override: the body is not welded to the world (move it through qpos)");
  CUDA_TRY(cudaSetDevice(s->device));
  return s->precision == B2S_F32 ? body_pose_override_t<float>(s, s->sf, body_id) : body_pose_override_t<double>(s, s->sd, body_id);
}

int b2s_env_step(b2s_sim* s, const void* action, int nsub) {
  if (!s || !s->has_ctrl || !action || nsub < 1) return fail(B2S_ERR_ARG, "b2s_env_step: bad argument / controller not configured");
  if (s->mode == 2 && !s->export_env_step && !s->profile)
    return launch_unit(s, PH_STEP1 | PH_STEP2 | PH_CTRL | (s->has_obs ? PH_OBS : 0), nsub, action);
  if (s->mode == 1 && !s->export_env_step && !s->profile)
    return launch_pipeline(s, PH_STEP1 | PH_STEP2 | PH_CTRL | (s->has_obs ? PH_OBS : 0), nsub, action);
  return launch(s, PH_STEP1 | PH_STEP2 | PH_CTRL | (s->has_obs ? PH_OBS : 0) | (s->export_env_step ? PH_EXPORT : 0) | (s->profile ? PH_PROFILE : 0), nsub, action);
}

int b2s_obs_config(b2s_sim* s, int obs_dim, const int* op, const int* a, const int* b) {
  if (!s || obs_dim <= 0 || !op || !a || !b) return fail(B2S_ERR_ARG, "b2s_obs_config: bad argument");
  CUDA_TRY(cudaSetDevice(s->device));
  try {
    std::vector<int> vo(op, op + obs_dim), va(a, a + obs_dim), vb(b, b + obs_dim);
    s->ctrl.obs_dim = obs_dim;
    s->ctrl.obs_op = dev_upload(s, vo); s->ctrl.obs_a = dev_upload(s, va); s->ctrl.obs_b = dev_upload(s, vb);
    if (obs_dim > 128) throw std::string("obs_dim > 128 not supported");
    int* fresh = state_arr_i(s, "obs_fresh", 0);
    {
      std::vector<int> ones(s->n_env, 1);
      if (cudaMemcpy(fresh, ones.data(), sizeof(int) * s->n_env, cudaMemcpyHostToDevice) != cudaSuccess) throw std::string("obs_fresh upload failed");
      cudaStreamSynchronize(cudaStreamLegacy);
    }
    if (s->precision == B2S_F32) { s->sf.obs = state_arr<float>(s, "obs", obs_dim); s->sf.task_out = state_arr<float>(s, "task_out", 8); s->sf.obs_fresh = fresh; }
    else { s->sd.obs = state_arr<double>(s, "obs", obs_dim); s->sd.task_out = state_arr<double>(s, "task_out", 8); s->sd.obs_fresh = fresh; }
  } catch (const std::string& e) { return fail(B2S_ERR_CUDA, e); }
  s->has_obs = 1;
  s->dirty = 1;
  return B2S_OK;
}

int b2s_task_config2(b2s_sim* s, int body2, const int* obj2, int no2) {
  if (!s || body2 >= s->nbody) return fail(B2S_ERR_ARG, "b2s_task_config2: bad argument");
  unsigned long long mk = 0;
  for (int i = 0; i < no2; i++) {
    if (obj2[i] < 0 || obj2[i] >= s->ngeom) return fail(B2S_ERR_ARG, "b2s_task_config2: geom id out of range");
    int k = s->cgid[obj2[i]];
    if (k >= 0) mk |= 1ull << k;
  }
  s->ctrl.task_body2 = body2; s->ctrl.mask_obj2 = mk; s->dirty = 1;
  return B2S_OK;
}

int b2s_task_table(b2s_sim* s, int n, const int* op, const int* a, const int* b) {
  if (!s || n <= 0 || n > 64 || !op || !a || !b) return fail(B2S_ERR_ARG, "b2s_task_table: bad argument");
  CUDA_TRY(cudaSetDevice(s->device));
  try {
    std::vector<int> vo(op, op + n), va(a, a + n), vb(b, b + n);
    s->ctrl.task_dim = n;
    s->ctrl.task_op = dev_upload(s, vo); s->ctrl.task_a = dev_upload(s, va); s->ctrl.task_b = dev_upload(s, vb);
    if (s->precision == B2S_F32) s->sf.task_vec = state_arr<float>(s, "task_vec", n);
    else s->sd.task_vec = state_arr<double>(s, "task_vec", n);
  } catch (const std::string& e) { return fail(B2S_ERR_CUDA, e); }
  s->dirty = 1;
  return B2S_OK;
}

int b2s_task_objects(b2s_sim* s, int nobjects, const int* geoms, const int* counts) {
  if (!s || nobjects < 0 || nobjects > 4 || (nobjects > 0 && (!geoms || !counts))) return fail(B2S_ERR_ARG, "b2s_task_objects: bad argument");
  int o = 0;
  for (int i = 0; i < 4; i++) s->ctrl.mask_objs[i] = 0;
  for (int i = 0; i < nobjects; i++)
    for (int k = 0; k < counts[i]; k++, o++) {
      if (geoms[o] < 0 || geoms[o] >= s->ngeom) return fail(B2S_ERR_ARG, "b2s_task_objects: geom id out of range");
      int cg = s->cgid[geoms[o]];
      if (cg >= 0) s->ctrl.mask_objs[i] |= 1ull << cg;
    }
  s->ctrl.n_objs = nobjects; s->dirty = 1;
  return B2S_OK;
}

int b2s_task_config(b2s_sim* s, int body, int site, const int* left, int nl, const int* right, int nr, const int* obj, int no) {
  if (!s || body < 0 || body >= s->nbody || site < 0 || site >= s->nsite) return fail(B2S_ERR_ARG, "b2s_task_config: bad argument");
  auto mk = [&](const int* g, int n, unsigned long long& out) {
    out = 0;
    for (int i = 0; i < n; i++) {
      if (g[i] < 0 || g[i] >= s->ngeom) return false;
      int k = s->cgid[g[i]];
      if (k >= 0) out |= 1ull << k;
    }
    return true;
  };
  if (!mk(left, nl, s->ctrl.mask_left) || !mk(right, nr, s->ctrl.mask_right) || !mk(obj, no, s->ctrl.mask_obj))
    return fail(B2S_ERR_ARG, "b2s_task_config: geom id out of range");
  s->ctrl.task_body = body; s->ctrl.task_site = site;
  if (s->ctrl.mask_obj2 == 0) s->ctrl.task_body2 = -1;
  s->dirty = 1;
  return B2S_OK;
}

}  // extern "C"
