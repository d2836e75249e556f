// capi.cu -- the extern "C" boundary of libb200sp (include/b200sp.h).  Catches every C++ exception and
// turns it into an int error code (PetscErrorCode convention); no C++ type crosses the boundary.
#include "solver.h"
#include "dist.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <sstream>

using namespace b200sp;

struct b200sp_ctx_s { Ctx c; };
struct b200sp_vec_s { Vec v; b200sp_vec_s(Ctx *c, int64_t n) : v(c, n) {} };
struct b200sp_mat_s {
  Mat m;
  std::shared_ptr<Csr> t; // explicit transpose cached by b200sp_mat_mult_transpose, valid for the value state t_state
  int64_t t_state = -1;
};
struct b200sp_ksp_s { Solver s; explicit b200sp_ksp_s(Ctx *c) : s(c) {} };
struct b200sp_pc_s { Solver s; explicit b200sp_pc_s(Ctx *c) : s(c) {} }; // a PC is the preconditioner half of the same solver object
struct b200sp_dmda_s { Dmda d; };
struct b200sp_dmda3d_s { Dmda3 d; };

static thread_local std::string g_last_error;

// One host thread may drive contexts on several devices (single-process multi-GPU): every entry point that launches
// work first makes its context's device current.
static inline void use_device(Ctx *c) {
  int cur = -1;
  if (cudaGetDevice(&cur) != cudaSuccess || cur != c->device) B2_CUDA(cudaSetDevice(c->device));
}
// device-side waits (peer-to-peer halos, small collectives) report a timeout through the context's mapped error word;
// it is checked -- and cleared, so that one failure does not poison every later call -- after every solve and PC apply
static inline void check_device_error(Ctx *c) {
  if (c->h_err && *c->h_err) {
    const int code = *c->h_err;
    *c->h_err = 0;
    throw Error(B200SP_ERR_NCCL, "peer-to-peer exchange timed out waiting for a neighbour (code " + std::to_string(code) + "); results of this call are invalid");
  }
}

#define API_BEGIN try {
#define API_END                                                    \
  return B200SP_OK;                                                \
  }                                                                \
  catch (const Error &e) { g_last_error = e.what(); return e.code; } \
  catch (const std::exception &e) { g_last_error = e.what(); return B200SP_ERR_ARG; } \
  catch (...) { g_last_error = "unknown error"; return B200SP_ERR_ARG; }

namespace b200sp {
Ctx::~Ctx() {
  delete reducer;
  delete dcomm;
  if (comm) nccl().CommDestroy(comm);
  if (d_partials) cudaFree(d_partials);
  if (d_ticket) cudaFree(d_ticket);
  if (d_scalars) cudaFree(d_scalars);
  if (h_scalars) cudaFreeHost(h_scalars);
  if (h_err) cudaFreeHost(h_err);
  if (pev0) cudaEventDestroy(pev0);
  if (pev1) cudaEventDestroy(pev1);
  if (tev0) cudaEventDestroy(tev0);
  if (tev1) cudaEventDestroy(tev1);
  if (stream) cudaStreamDestroy(stream);
  if (stream2) cudaStreamDestroy(stream2);
}

// ---- host-only DMDA index arithmetic ----
void dmda_proc_grid(int M, int N, int size, int *pm, int *pn) {
  int m = (int)(0.5 + std::sqrt(((double)M) * ((double)size) / ((double)N))), n = 1;
  if (!m) m = 1;
  while (m > 0) {
    n = size / m;
    if (m * n == size) break;
    m--;
  }
  if (M > N && m < n) std::swap(m, n);
  *pm = m;
  *pn = n;
}
void dmda_ownership(int M, int m, int *lx) {
  for (int i = 0; i < m; ++i) lx[i] = M / m + ((M % m) > i);
}
} // namespace b200sp

extern "C" {

const char *b200sp_last_error(void) { return g_last_error.c_str(); }
const char *b200sp_version(void) { return "b200sp 0.1 (sm_100a)"; }

int b200sp_nccl_unique_id(char id[128]) {
  API_BEGIN
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId u;
  if (!nccl().ok) throw Error(B200SP_ERR_NCCL, "NCCL unavailable: " + nccl().err);
  B2_NCCL(nccl().GetUniqueId(&u));
  std::memcpy(id, &u, 128);
  API_END
}

int b200sp_ctx_create(int device, int rank, int size, const char nccl_id[128], b200sp_ctx *out) {
  API_BEGIN
  B2_REQUIRE(out && size >= 1 && rank >= 0 && rank < size, "ctx_create: bad arguments");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    throw Error(B200SP_ERR_NO_DEVICE, "no CUDA device: libb200sp has no CPU fallback");
  }
  B2_REQUIRE(device >= 0 && device < ndev, "ctx_create: bad device index");
  B2_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  B2_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) throw Error(B200SP_ERR_NO_DEVICE, std::string("device ") + prop.name + " is not sm_100: libb200sp is built for B200 only");
  auto *h = new b200sp_ctx_s();
  Ctx &c = h->c;
  try {
    c.device = device; c.rank = rank; c.size = size;
    c.num_sms = prop.multiProcessorCount;
    B2_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    B2_CUDA(cudaStreamCreateWithFlags(&c.stream2, cudaStreamNonBlocking));
    B2_CUDA(cudaMalloc((void **)&c.d_partials, sizeof(double) * RED_MAX_BLOCKS * RED_MAX_OUT));
    B2_CUDA(cudaMalloc((void **)&c.d_ticket, sizeof(unsigned)));
    B2_CUDA(cudaMemset(c.d_ticket, 0, sizeof(unsigned)));
    B2_CUDA(cudaMalloc((void **)&c.d_scalars, sizeof(double) * N_SCALARS));
    B2_CUDA(cudaMemset(c.d_scalars, 0, sizeof(double) * N_SCALARS));
    B2_CUDA(cudaMallocHost((void **)&c.h_scalars, sizeof(double) * N_SCALARS));
    B2_CUDA(cudaHostAlloc((void **)&c.h_err, sizeof(int), cudaHostAllocMapped));
    *c.h_err = 0;
    B2_CUDA(cudaHostGetDevicePointer((void **)&c.d_err, c.h_err, 0));
    B2_CUDA(cudaEventCreate(&c.pev0)); B2_CUDA(cudaEventCreate(&c.pev1));
    B2_CUDA(cudaEventCreate(&c.tev0)); B2_CUDA(cudaEventCreate(&c.tev1));
    if (size > 1) {
      B2_REQUIRE(nccl_id, "ctx_create: size > 1 needs an NCCL unique id");
      ncclUniqueId u;
      std::memcpy(&u, nccl_id, 128);
      if (!nccl().ok) throw Error(B200SP_ERR_NCCL, "NCCL unavailable: " + nccl().err);
      B2_NCCL(nccl().CommInitRank(&c.comm, size, u, rank));
      c.dcomm = make_nccl_comm(c.comm, rank, size);
      c.reducer = make_collective(&c, 32);
    }
  } catch (...) { delete h; throw; }
  *out = h;
  API_END
}
// ---- in-process rank group: all ranks are threads of this process (tests on a 1-GPU box; single-process multi-GPU)
struct b200sp_group_s { std::shared_ptr<LocalGroup> g; };
int b200sp_local_group_create(int size, b200sp_group *out) {
  API_BEGIN
  B2_REQUIRE(out && size >= 1, "local_group_create: bad arguments");
  auto *h = new b200sp_group_s();
  h->g = std::make_shared<LocalGroup>(size);
  *out = h;
  API_END
}
int b200sp_local_group_destroy(b200sp_group g) { API_BEGIN delete g; API_END }
int b200sp_ctx_create_local(b200sp_group group, int rank, int device, b200sp_ctx *out) {
  API_BEGIN
  B2_REQUIRE(group && out && rank >= 0 && rank < group->g->size, "ctx_create_local: bad arguments");
  b200sp_ctx ctx = nullptr;
  int rc = b200sp_ctx_create(device, 0, 1, nullptr, &ctx);
  if (rc) throw Error(rc, g_last_error);
  ctx->c.rank = rank;
  ctx->c.size = group->g->size;
  if (group->g->size > 1) { ctx->c.dcomm = make_local_comm(group->g, rank, device); ctx->c.reducer = make_collective(&ctx->c, 32); }
  *out = ctx;
  API_END
}
int b200sp_ctx_destroy(b200sp_ctx ctx) {
  API_BEGIN
  if (ctx) { cudaSetDevice(ctx->c.device); cudaStreamSynchronize(ctx->c.stream); delete ctx; }
  API_END
}
int b200sp_ctx_synchronize(b200sp_ctx ctx) { API_BEGIN ctx->c.sync(); API_END }
int b200sp_ctx_get_stream(b200sp_ctx ctx, void **s) { API_BEGIN *s = (void *)ctx->c.stream; API_END }
int b200sp_ctx_get_launch_count(b200sp_ctx ctx, int64_t *count) { API_BEGIN *count = ctx->c.launches; API_END }
int b200sp_ctx_timer_start(b200sp_ctx ctx) { API_BEGIN B2_CUDA(cudaEventRecord(ctx->c.tev0, ctx->c.stream)); API_END }
int b200sp_ctx_timer_stop(b200sp_ctx ctx, double *ms) {
  API_BEGIN
  B2_CUDA(cudaEventRecord(ctx->c.tev1, ctx->c.stream));
  B2_CUDA(cudaEventSynchronize(ctx->c.tev1));
  float f = 0;
  B2_CUDA(cudaEventElapsedTime(&f, ctx->c.tev0, ctx->c.tev1));
  *ms = f;
  API_END
}
int b200sp_ctx_profile_enable(b200sp_ctx ctx, int on) {
  API_BEGIN
  ctx->c.sync();
  ctx->c.profile = on != 0;
  if (on) ctx->c.prof.clear();
  API_END
}
int b200sp_ctx_profile_report(b200sp_ctx ctx, char *buf, int buflen) {
  API_BEGIN
  std::ostringstream o;
  o << "{";
  bool first = true;
  for (auto &kv : ctx->c.prof) {
    if (!first) o << ", ";
    first = false;
    o << "\"" << kv.first << "\": {\"ms\": " << kv.second.ms << ", \"launches\": " << kv.second.n << "}";
  }
  o << "}";
  std::string s = o.str();
  B2_REQUIRE(buf && buflen > (int)s.size(), "profile_report: buffer too small");
  std::memcpy(buf, s.c_str(), s.size() + 1);
  API_END
}

// ---------------------------------------------------------------- DMDA
int b200sp_dmda_proc_grid(int M, int N, int size, int *m, int *n) { API_BEGIN B2_REQUIRE(M > 0 && N > 0 && size > 0, "bad args"); dmda_proc_grid(M, N, size, m, n); API_END }
int b200sp_dmda_ownership(int M, int m, int *lx) { API_BEGIN B2_REQUIRE(M > 0 && m > 0 && lx, "bad args"); dmda_ownership(M, m, lx); API_END }
int b200sp_dmda_corners(int M, int N, int size, int rank, int *xs, int *ys, int *xm, int *ym) {
  API_BEGIN
  Layout L(M, N, size);
  B2_REQUIRE(rank >= 0 && rank < size, "bad rank");
  const int pi = rank % L.m, pj = rank / L.m;
  *xs = L.xoff[pi]; *ys = L.yoff[pj]; *xm = L.lx[pi]; *ym = L.ly[pj];
  API_END
}
int b200sp_dmda_element_corners(int M, int N, int size, int rank, int *si, int *sj, int *ni, int *nj) {
  API_BEGIN
  Layout L(M, N, size);
  B2_REQUIRE(rank >= 0 && rank < size, "bad rank");
  const int pi = rank % L.m, pj = rank / L.m;
  const int xs = L.xoff[pi], ys = L.yoff[pj];
  const int gxs = xs > 0 ? xs - 1 : xs, gys = ys > 0 ? ys - 1 : ys; // DMDAGetElementsCorners
  *si = gxs; *sj = gys;
  *ni = xs + L.lx[pi] - gxs - 1;
  *nj = ys + L.ly[pj] - gys - 1;
  API_END
}
int b200sp_dmda_global_node(int M, int N, int size, int i, int j, int *gnode, int *owner) {
  API_BEGIN
  Layout L(M, N, size);
  B2_REQUIRE(i >= 0 && i < M && j >= 0 && j < N, "node out of range");
  if (gnode) *gnode = L.gnode(i, j);
  if (owner) *owner = L.owner(i, j);
  API_END
}
int b200sp_dmda_halo_plan(int M, int N, int size, int rank, int *nghost, int *ghost_gnode, int *ghost_owner, int *nsend_total, int *send_rank,
                          int *send_lnode) {
  API_BEGIN
  Layout L(M, N, size);
  B2_REQUIRE(rank >= 0 && rank < size, "bad rank");
  const HaloPlan P = plan_halo(L, rank); // the same plan the device halo is built from
  if (nghost) *nghost = (int)P.ghost_gnode.size();
  if (ghost_gnode) std::copy(P.ghost_gnode.begin(), P.ghost_gnode.end(), ghost_gnode);
  if (ghost_owner) std::copy(P.ghost_owner.begin(), P.ghost_owner.end(), ghost_owner);
  if (nsend_total) *nsend_total = (int)P.send_lnode.size();
  if (send_rank)
    for (const HaloMsg &m : P.msgs)
      for (int64_t k = 0; k < m.send_cnt; ++k) send_rank[m.send_off + k] = m.peer;
  if (send_lnode) std::copy(P.send_lnode.begin(), P.send_lnode.end(), send_lnode);
  API_END
}
int b200sp_dmda_halo_push_table(int M, int N, int size, int rank, int *n_owned, int *node_ent, int *n_entries, int *entry_rank, int *entry_pos) {
  API_BEGIN
  Layout L(M, N, size);
  B2_REQUIRE(rank >= 0 && rank < size, "bad rank");
  const HaloPlan P = plan_halo(L, rank);
  B2_REQUIRE(P.push_valid, "halo push table: a node goes to more than 7 neighbours (boxes thinner than 2 nodes)");
  if (n_owned) *n_owned = P.n_owned;
  if (node_ent) std::copy(P.push_node_ent.begin(), P.push_node_ent.begin() + P.n_owned, node_ent);
  if (n_entries) *n_entries = (int)P.push_ent_msg.size();
  if (entry_rank) for (size_t e = 0; e < P.push_ent_msg.size(); ++e) entry_rank[e] = e ? P.msgs[(size_t)P.push_ent_msg[e]].peer : -1;
  if (entry_pos) std::copy(P.push_ent_pos.begin(), P.push_ent_pos.end(), entry_pos);
  API_END
}
int b200sp_dmda_create(b200sp_ctx ctx, int M, int N, b200sp_dmda *da) {
  API_BEGIN
  B2_REQUIRE(ctx && da, "dmda_create: bad arguments");
  use_device(&ctx->c);
  Layout L(M, N, ctx->c.size);
  auto *h = new b200sp_dmda_s();
  Dmda &d = h->d;
  d.ctx = &ctx->c; d.M = M; d.N = N; d.pm = L.m; d.pn = L.n; d.lx = L.lx; d.ly = L.ly;
  L.box(ctx->c.rank, &d.xs, &d.ys, &d.xm, &d.ym);
  if (ctx->c.size > 1) {
    try {
      d.layout = std::make_shared<Layout>(L);
      d.halo = make_halo(&ctx->c, L, ctx->c.rank);
    } catch (...) { delete h; throw; }
  }
  *da = h;
  API_END
}
int b200sp_dmda_destroy(b200sp_dmda da) { API_BEGIN delete da; API_END }
int b200sp_dmda_get_info(b200sp_dmda da, int *M, int *N, int *xs, int *ys, int *xm, int *ym) {
  API_BEGIN
  if (M) *M = da->d.M; if (N) *N = da->d.N;
  if (xs) *xs = da->d.xs; if (ys) *ys = da->d.ys; if (xm) *xm = da->d.xm; if (ym) *ym = da->d.ym;
  API_END
}
int b200sp_dmda_bc_ids(b200sp_dmda da, int dof, int *n, int *ids) {
  API_BEGIN
  std::vector<int> v = dmda_bc_ids(da->d, dof);
  if (n) *n = (int)v.size();
  if (ids) std::copy(v.begin(), v.end(), ids);
  API_END
}

// ---------------------------------------------------------------- Vec
int b200sp_vec_create(b200sp_ctx ctx, int64_t n, b200sp_vec *v) {
  API_BEGIN
  B2_REQUIRE(ctx && v && n >= 0, "vec_create: bad arguments");
  use_device(&ctx->c);
  *v = new b200sp_vec_s(&ctx->c, n);
  API_END
}
int b200sp_vec_destroy(b200sp_vec v) { API_BEGIN if (v) { v->v.ctx->sync(); delete v; } API_END }
int b200sp_vec_get_size(b200sp_vec v, int64_t *n) { API_BEGIN *n = v->v.n; API_END }
int b200sp_vec_set(b200sp_vec v, double a) { API_BEGIN vec_set(v->v.ctx, v->v.n, a, v->v.d); API_END }
int b200sp_vec_set_values_host(b200sp_vec v, int64_t n, const int *idx, const double *vals) {
  API_BEGIN
  Ctx *c = v->v.ctx;
  if (n > 0) {
    // INSERT_VALUES of a host list: staged through the device copy (values are few: boundary dofs)
    std::vector<double> h((size_t)v->v.n);
    B2_CUDA(cudaMemcpyAsync(h.data(), v->v.d, sizeof(double) * (size_t)v->v.n, cudaMemcpyDeviceToHost, c->stream));
    c->sync();
    for (int64_t t = 0; t < n; ++t) {
      B2_REQUIRE(idx[t] >= 0 && idx[t] < v->v.n, "vec_set_values: index out of range");
      h[(size_t)idx[t]] = vals[t];
    }
    B2_CUDA(cudaMemcpyAsync(v->v.d, h.data(), sizeof(double) * (size_t)v->v.n, cudaMemcpyHostToDevice, c->stream));
    c->sync();
  }
  API_END
}
int b200sp_vec_copy_from_host(b200sp_vec v, const double *host, int64_t n) {
  API_BEGIN
  B2_REQUIRE(n == v->v.n, "vec_copy_from_host: size mismatch");
  B2_CUDA(cudaMemcpyAsync(v->v.d, host, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, v->v.ctx->stream));
  v->v.ctx->sync();
  API_END
}
int b200sp_vec_copy_to_host(b200sp_vec v, double *host, int64_t n) {
  API_BEGIN
  B2_REQUIRE(n == v->v.n, "vec_copy_to_host: size mismatch");
  B2_CUDA(cudaMemcpyAsync(host, v->v.d, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, v->v.ctx->stream));
  v->v.ctx->sync();
  API_END
}
int b200sp_vec_get_device_ptr(b200sp_vec v, double **dev) { API_BEGIN *dev = v->v.d; API_END }
#define SAME_SIZE(a, b) B2_REQUIRE((a)->v.n == (b)->v.n, "vector size mismatch")
int b200sp_vec_copy(b200sp_vec x, b200sp_vec y) { API_BEGIN SAME_SIZE(x, y); vec_copy(x->v.ctx, x->v.n, x->v.d, y->v.d); API_END }
int b200sp_vec_scale(b200sp_vec x, double a) { API_BEGIN vec_scale(x->v.ctx, x->v.n, a, x->v.d); API_END }
int b200sp_vec_axpy(b200sp_vec y, double a, b200sp_vec x) { API_BEGIN SAME_SIZE(x, y); vec_axpy(y->v.ctx, y->v.n, a, x->v.d, y->v.d); API_END }
int b200sp_vec_aypx(b200sp_vec y, double a, b200sp_vec x) { API_BEGIN SAME_SIZE(x, y); vec_aypx(y->v.ctx, y->v.n, a, x->v.d, y->v.d); API_END }
int b200sp_vec_waxpy(b200sp_vec w, double a, b200sp_vec x, b200sp_vec y) {
  API_BEGIN SAME_SIZE(x, y); SAME_SIZE(x, w); vec_waxpy(w->v.ctx, w->v.n, a, x->v.d, y->v.d, w->v.d); API_END
}
int b200sp_vec_pointwise_mult(b200sp_vec w, b200sp_vec x, b200sp_vec y) {
  API_BEGIN SAME_SIZE(x, y); SAME_SIZE(x, w); vec_pointwise_mult(w->v.ctx, w->v.n, x->v.d, y->v.d, w->v.d); API_END
}
static void global_sum(Ctx *c, int k, double *host) {
  if (c->reducer) c->reducer->allreduce_sum(c->d_scalars, k, c->stream);
  c->fetch_scalars(c->d_scalars, k, host);
}
int b200sp_vec_dot(b200sp_vec x, b200sp_vec y, double *result) {
  API_BEGIN
  SAME_SIZE(x, y);
  Ctx *c = x->v.ctx;
  vec_dot(c, x->v.n, x->v.d, y->v.d, c->d_scalars);
  global_sum(c, 1, result);
  API_END
}
int b200sp_vec_norm(b200sp_vec x, double *result) {
  API_BEGIN
  Ctx *c = x->v.ctx;
  vec_dot(c, x->v.n, x->v.d, x->v.d, c->d_scalars);
  global_sum(c, 1, result);
  *result = std::sqrt(*result);
  API_END
}
int b200sp_vec_mdot(b200sp_vec x, int k, const b200sp_vec *y, double *result) {
  API_BEGIN
  B2_REQUIRE(k >= 0 && k <= N_SCALARS, "vec_mdot: k out of range");
  Ctx *c = x->v.ctx;
  // the y's are independent allocations: one fused launch per group is only possible for a strided basis
  // (the KSP path); here each dot is its own launch into consecutive result slots.
  for (int j = 0; j < k; ++j) { SAME_SIZE(x, y[j]); vec_dot(c, x->v.n, x->v.d, y[j]->v.d, c->d_scalars + j); }
  if (k) global_sum(c, k, result);
  API_END
}
int b200sp_vec_maxpy(b200sp_vec y, int k, const double *a, const b200sp_vec *x) {
  API_BEGIN
  for (int j = 0; j < k; ++j) { SAME_SIZE(y, x[j]); vec_axpy(y->v.ctx, y->v.n, a[j], x[j]->v.d, y->v.d); }
  API_END
}

// measurement hook (bench.py --config sweep): the fused Gram-Schmidt kernels of the GMRES drivers -- VecMDot of w against a
// strided basis of k vectors, and VecMAXPY fused with the norm -- timed with CUDA events over `reps` launches each
int b200sp_bench_orthogonalization(b200sp_ctx ctx, int64_t n, int k, int reps, double *ms_mdot, double *ms_maxpy) {
  API_BEGIN
  B2_REQUIRE(ctx && n > 0 && k >= 1 && k <= 30 && reps >= 1, "bench_orthogonalization: bad arguments");
  Ctx *c = &ctx->c;
  use_device(c);
  const int64_t ld = (n + 15) & ~(int64_t)15;
  DevBuf<double> V((size_t)ld * k), w((size_t)ld);
  vec_hash(c, ld * k, V.p);
  vec_hash(c, n, w.p);
  double *h = c->d_scalars;
  auto timed = [&](auto fn) {
    fn();
    B2_CUDA(cudaEventRecord(c->tev0, c->stream));
    for (int r = 0; r < reps; ++r) fn();
    B2_CUDA(cudaEventRecord(c->tev1, c->stream));
    B2_CUDA(cudaEventSynchronize(c->tev1));
    float f = 0;
    B2_CUDA(cudaEventElapsedTime(&f, c->tev0, c->tev1));
    return (double)f / reps;
  };
  const double t1 = timed([&] { vec_mdot(c, n, k, w.p, V.p, ld, h); });
  B2_CUDA(cudaMemsetAsync(h, 0, sizeof(double) * (size_t)(k + 1), c->stream)); // zero coefficients: w stays bounded over the repetitions
  const double t2 = timed([&] { vec_maxpy_norm2(c, n, k, w.p, V.p, ld, h, h + k); });
  if (ms_mdot) *ms_mdot = t1;
  if (ms_maxpy) *ms_maxpy = t2;
  c->sync();
  API_END
}

// ---------------------------------------------------------------- Mat
static b200sp_mat wrap(Ctx *c, std::shared_ptr<Csr> A) {
  auto *h = new b200sp_mat_s();
  h->m.ctx = c;
  h->m.csr = A;
  return h;
}
static Csr &plain(b200sp_mat A) {
  B2_REQUIRE(A && !A->m.nest && A->m.csr, "operation needs a plain CSR matrix, not a nest");
  return *A->m.csr;
}
int b200sp_mat_create_csr(b200sp_ctx ctx, int nrows, int ncols, const int *rowptr, const int *col, const double *val, b200sp_mat *A) {
  API_BEGIN *A = wrap(&ctx->c, csr_from_host(&ctx->c, nrows, ncols, rowptr, col, val)); API_END
}
int b200sp_mat_create_coo(b200sp_ctx ctx, int nrows, int ncols, int64_t ncoo, const int *row, const int *col, const double *val, b200sp_mat *A) {
  API_BEGIN *A = wrap(&ctx->c, csr_from_coo_host(&ctx->c, nrows, ncols, ncoo, row, col, val)); API_END
}
int b200sp_mat_destroy(b200sp_mat A) { API_BEGIN if (A) { A->m.ctx->sync(); delete A; } API_END }
int b200sp_mat_set_grid(b200sp_mat A, int M, int N, int dof) {
  API_BEGIN
  Csr &m = plain(A);
  B2_REQUIRE(M >= 2 && N >= 2 && dof >= 1 && (int64_t)M * N * dof == m.nrows && m.nrows == m.ncols, "mat_set_grid: grid does not match the matrix");
  m.grid_M = M; m.grid_N = N; m.dof_r = dof; m.dof_c = dof;
  API_END
}
int b200sp_mat_get_size(b200sp_mat A, int *nrows, int *ncols, int64_t *nnz) {
  API_BEGIN
  if (nrows) *nrows = A->m.nrows();
  if (ncols) *ncols = A->m.ncols();
  if (nnz) {
    if (A->m.nest) { *nnz = 0; for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) if (A->m.blk[i][j]) *nnz += A->m.blk[i][j]->nnz; }
    else *nnz = A->m.csr->nnz;
  }
  API_END
}
int b200sp_mat_get_csr_host(b200sp_mat A, int *rowptr, int *col, double *val) {
  API_BEGIN
  Csr &M = plain(A);
  Ctx *c = M.ctx;
  if (M.halo) {
    // row-partitioned matrix: local rows with GLOBAL (PETSc numbering) column ids, sorted -- what MatView / MatGetRow
    // show for an MPIAIJ matrix
    std::vector<int> rp((size_t)M.nrows + 1), cj((size_t)M.nnz + 1);
    std::vector<double> va((size_t)M.nnz + 1);
    B2_CUDA(cudaMemcpyAsync(rp.data(), M.rowptr.p, sizeof(int) * ((size_t)M.nrows + 1), cudaMemcpyDeviceToHost, c->stream));
    if (M.nnz) B2_CUDA(cudaMemcpyAsync(cj.data(), M.col.p, sizeof(int) * (size_t)M.nnz, cudaMemcpyDeviceToHost, c->stream));
    if (M.nnz) B2_CUDA(cudaMemcpyAsync(va.data(), M.val.p, sizeof(double) * (size_t)M.nnz, cudaMemcpyDeviceToHost, c->stream));
    c->sync();
    const int dofc = M.halo_dof;
    std::vector<std::pair<int, double>> row;
    for (int r = 0; r < M.nrows; ++r) {
      if (rowptr) rowptr[r] = rp[(size_t)r];
      row.clear();
      for (int k = rp[(size_t)r]; k < rp[(size_t)r + 1]; ++k) {
        const int cl = cj[(size_t)k];
        const int g = cl < M.ncols ? (int)(M.col_gstart + cl) : M.halo->ghost_gnode[(size_t)((cl - M.ncols) / dofc)] * dofc + (cl - M.ncols) % dofc;
        row.push_back({g, va[(size_t)k]});
      }
      std::sort(row.begin(), row.end(), [](const std::pair<int, double> &a, const std::pair<int, double> &b) { return a.first < b.first; });
      int64_t p = rp[(size_t)r];
      for (auto &e : row) { if (col) col[p] = e.first; if (val) val[p] = e.second; ++p; }
    }
    if (rowptr) rowptr[M.nrows] = rp[(size_t)M.nrows];
    return B200SP_OK;
  }
  if (rowptr) B2_CUDA(cudaMemcpyAsync(rowptr, M.rowptr.p, sizeof(int) * ((size_t)M.nrows + 1), cudaMemcpyDeviceToHost, c->stream));
  if (col && M.nnz) B2_CUDA(cudaMemcpyAsync(col, M.col.p, sizeof(int) * (size_t)M.nnz, cudaMemcpyDeviceToHost, c->stream));
  if (val && M.nnz) B2_CUDA(cudaMemcpyAsync(val, M.val.p, sizeof(double) * (size_t)M.nnz, cudaMemcpyDeviceToHost, c->stream));
  c->sync();
  API_END
}
int b200sp_mat_get_spmv_plan(b200sp_mat A, int64_t hist[14], int *kernel, int *max_row_nnz) {
  API_BEGIN
  Csr &M = plain(A);
  if (hist) std::copy(M.hist, M.hist + 14, hist);
  if (kernel) *kernel = M.kernel;
  if (max_row_nnz) *max_row_nnz = M.max_row_nnz;
  API_END
}
int b200sp_mat_get_spmv_format(b200sp_mat A, int *block_r, int *block_c, int *value_dict, int64_t *matrix_bytes) {
  API_BEGIN
  Csr &M = plain(A);
  const bool blk = M.bcol.p != nullptr, dict = M.dict_state == 1;
  if (block_r) *block_r = dict ? M.pd_br : blk ? M.blk_r : 1;
  if (block_c) *block_c = dict ? M.pd_bc : blk ? M.blk_c : 1;
  if (value_dict) *value_dict = dict ? 1 : 0;
  if (matrix_bytes) {
    const int br = blk ? M.blk_r : 1, bc = blk ? M.blk_c : 1;
    int64_t bytes;
    if (dict) bytes = M.dict_bytes; // tile blobs (values, patterns, codes) + the tile offsets: nothing else is read
    else bytes = 8 * M.nnz + 4 * (M.nnz / (br * bc)) + 4 * (int64_t)(M.nrows + 1) + (blk ? 4 * (int64_t)(M.nrows / br + 1) : 0);
    *matrix_bytes = bytes;
  }
  API_END
}
int b200sp_mat_set_spmv_format(b200sp_mat A, int block_index, int value_dict) {
  API_BEGIN
  Csr &M = plain(A);
  use_device(M.ctx);
  M.ctx->sync();
  csr_drop_value_dict(M);
  M.state++; // derived storage changed: a KSP that recorded CUDA graphs on the old storage sets up again at its next solve
  M.no_value_dict = value_dict == 0;
  if (!block_index) { M.bcol.release(); M.bptr.release(); M.blk_r = M.blk_c = 1; }
  else if (!M.bcol.p && M.dof_r > 0 && M.dof_c > 0 && M.dof_r * M.dof_c > 1) csr_try_block_index(M, M.dof_r, M.dof_c);
  API_END
}
int b200sp_mat_set_spmv_kernel(b200sp_mat A, int kernel) {
  API_BEGIN
  Csr &M = plain(A);
  B2_REQUIRE(kernel >= 0 && kernel <= 3, "bad kernel id");
  B2_REQUIRE((kernel != SPMV_STREAM && kernel != SPMV_TMA) || M.max_group_nnz <= 1152, "stream kernel: rows too long for the shared tile");
  M.kernel = kernel;
  API_END
}
static void mat_apply(Mat &m, const double *x, double *y, double alpha, const double *z, double beta_z) {
  use_device(m.ctx);
  if (!m.nest) { csr_spmv(*m.csr, x, y, alpha, z, beta_z); return; }
  const int64_t n0c = m.blk[0][0]->ncols, n0r = m.blk[0][0]->nrows;
  csr_spmv(*m.blk[0][0], x, y, alpha, z, beta_z);
  csr_spmv(*m.blk[0][1], x + n0c, y, alpha, y, 1.0);
  csr_spmv(*m.blk[1][0], x, y + n0r, alpha, z ? z + n0r : nullptr, beta_z);
  if (m.blk[1][1]) csr_spmv(*m.blk[1][1], x + n0c, y + n0r, alpha, y + n0r, 1.0);
}
int b200sp_mat_mult(b200sp_mat A, b200sp_vec x, b200sp_vec y) {
  API_BEGIN
  B2_REQUIRE(x->v.n == A->m.ncols() && y->v.n == A->m.nrows() && x != y, "mat_mult: size mismatch or aliasing");
  mat_apply(A->m, x->v.d, y->v.d, 1.0, nullptr, 0.0);
  API_END
}
int b200sp_mat_mult_add(b200sp_mat A, b200sp_vec x, b200sp_vec y, b200sp_vec z) {
  API_BEGIN
  B2_REQUIRE(x->v.n == A->m.ncols() && y->v.n == A->m.nrows() && z->v.n == y->v.n && x != z, "mat_mult_add: size mismatch or aliasing");
  mat_apply(A->m, x->v.d, z->v.d, 1.0, y->v.d, 1.0);
  API_END
}
int b200sp_mat_residual(b200sp_mat A, b200sp_vec b, b200sp_vec x, b200sp_vec r) {
  API_BEGIN
  B2_REQUIRE(x->v.n == A->m.ncols() && b->v.n == A->m.nrows() && r->v.n == b->v.n && x != r, "mat_residual: size mismatch or aliasing");
  mat_apply(A->m, x->v.d, r->v.d, -1.0, b->v.d, 1.0);
  API_END
}
int b200sp_mat_get_diagonal(b200sp_mat A, b200sp_vec d) {
  API_BEGIN
  Csr &M = plain(A);
  B2_REQUIRE(d->v.n == M.nrows, "mat_get_diagonal: size mismatch");
  csr_get_diagonal(M, d->v.d);
  API_END
}
int b200sp_mat_mult_transpose(b200sp_mat A, b200sp_vec x, b200sp_vec y) {
  API_BEGIN
  Csr &M = plain(A);
  B2_REQUIRE(!M.halo, "MatMultTranspose: not available for row-partitioned matrices (assemble the transpose explicitly)");
  B2_REQUIRE(x->v.n == M.nrows && y->v.n == M.ncols && x != y, "MatMultTranspose: size mismatch or aliasing");
  // the transpose is built once per value state by the stable device sort; its rows list the entries of a column of A by
  // ascending row, i.e. the order in which MatMultTranspose_SeqAIJ adds them -> same bits as the sequential scatter loop
  if (!A->t || A->t_state != M.state) { A->t = csr_transpose(M); A->t_state = M.state; }
  csr_spmv(*A->t, x->v.d, y->v.d);
  API_END
}
int b200sp_mat_transpose(b200sp_mat A, b200sp_mat *At) { API_BEGIN *At = wrap(A->m.ctx, csr_transpose(plain(A))); API_END }
int b200sp_mat_matmult(b200sp_mat A, b200sp_mat B, b200sp_mat *C) { API_BEGIN *C = wrap(A->m.ctx, csr_matmat(plain(A), plain(B))); API_END }
int b200sp_mat_scale_columns(b200sp_mat A, b200sp_vec d, b200sp_mat *C) {
  API_BEGIN
  Csr &M = plain(A);
  use_device(M.ctx);
  B2_REQUIRE(d->v.n == M.ncols, "mat_scale_columns: vector does not match the (owned) column space");
  *C = wrap(M.ctx, csr_scale_cols(M, d->v.d));
  API_END
}
int b200sp_mat_add_scaled(b200sp_mat A, double s, b200sp_mat B, b200sp_mat *C) {
  API_BEGIN
  use_device(A->m.ctx);
  *C = wrap(A->m.ctx, csr_add_scaled(plain(A), s, plain(B)));
  API_END
}
int b200sp_amg_aggregate(b200sp_mat A, int bs, double theta, int order, int *agg_host, int *nagg) {
  API_BEGIN
  Csr &M = plain(A);
  use_device(M.ctx);
  DevBuf<int> agg;
  *nagg = amg_aggregate(M, bs, theta, order, agg);
  if (agg_host && M.nrows / bs > 0) B2_CUDA(cudaMemcpy(agg_host, agg.p, sizeof(int) * (size_t)(M.nrows / bs), cudaMemcpyDeviceToHost));
  API_END
}
int b200sp_amg_prolongator(b200sp_mat A, int bs, double theta, int order, double omega, const int *node_weight, int *coarse_weight, b200sp_mat *P) {
  API_BEGIN
  Csr &M = plain(A);
  use_device(M.ctx);
  DevBuf<int> agg, w, wc;
  const int nagg = amg_aggregate(M, bs, theta, order, agg);
  const int nn = M.nrows / bs;
  if (node_weight && nn > 0) {
    for (int i = 0; i < nn; ++i) B2_REQUIRE(node_weight[i] >= 1, "amg_prolongator: node weights must be positive");
    w.alloc((size_t)nn);
    B2_CUDA(cudaMemcpy(w.p, node_weight, sizeof(int) * (size_t)nn, cudaMemcpyHostToDevice));
  }
  auto Pt = amg_tentative(M.ctx, nn, bs, agg, nagg, w.p, wc);
  if (coarse_weight && nagg > 0) B2_CUDA(cudaMemcpy(coarse_weight, wc.p, sizeof(int) * (size_t)nagg, cudaMemcpyDeviceToHost));
  *P = wrap(M.ctx, omega != 0.0 ? amg_smooth_prolongator(M, *Pt, omega) : Pt);
  API_END
}
int b200sp_mat_zero_rows_columns(b200sp_mat A, int n, const int *rows, double diag) {
  API_BEGIN
  Csr &M = plain(A);
  B2_REQUIRE(M.nrows == M.ncols, "MatZeroRowsColumns: matrix must be square");
  for (int t = 0; t < n; ++t) B2_REQUIRE(rows[t] >= 0 && rows[t] < M.nrows, "MatZeroRowsColumns: row out of range");
  csr_zero_rows_cols(M, n, rows, diag, true, true, true);
  API_END
}
int b200sp_mat_zero_rows(b200sp_mat A, int n, const int *rows, double diag) {
  API_BEGIN
  Csr &M = plain(A);
  for (int t = 0; t < n; ++t) B2_REQUIRE(rows[t] >= 0 && rows[t] < M.nrows, "MatZeroRows: row out of range");
  csr_zero_rows_cols(M, n, rows, diag, true, false, M.nrows == M.ncols && diag != 0.0);
  API_END
}
int b200sp_mat_zero_columns(b200sp_mat A, int n, const int *cols) {
  API_BEGIN
  Csr &M = plain(A);
  for (int t = 0; t < n; ++t) B2_REQUIRE(cols[t] >= 0 && cols[t] < M.ncols, "zero_columns: column out of range");
  csr_zero_rows_cols(M, n, cols, 0.0, false, true, false);
  API_END
}
int b200sp_mat_create_nest(b200sp_mat A00, b200sp_mat A01, b200sp_mat A10, b200sp_mat A11, b200sp_mat *K) {
  API_BEGIN
  Csr &a00 = plain(A00), &a01 = plain(A01), &a10 = plain(A10);
  B2_REQUIRE(a00.nrows == a01.nrows && a00.ncols == a10.ncols, "nest: block shapes do not match");
  if (A11) B2_REQUIRE(plain(A11).nrows == a10.nrows && plain(A11).ncols == a01.ncols, "nest: A11 shape does not match");
  auto *h = new b200sp_mat_s();
  h->m.ctx = a00.ctx;
  h->m.nest = true;
  h->m.blk[0][0] = A00->m.csr; h->m.blk[0][1] = A01->m.csr; h->m.blk[1][0] = A10->m.csr;
  h->m.blk[1][1] = A11 ? A11->m.csr : nullptr;
  *K = h;
  API_END
}

// ---------------------------------------------------------------- assembly
int b200sp_assemble_stress_coeff(b200sp_dmda da, int coeff_kind, b200sp_mat *A) {
  API_BEGIN
  B2_REQUIRE(da && A && (coeff_kind == 0 || coeff_kind == 1), "assemble_stress_coeff: bad arguments");
  use_device(da->d.ctx);
  *A = wrap(da->d.ctx, assemble_stress(da->d, 0, coeff_kind));
  API_END
}
int b200sp_assemble_stress(b200sp_dmda da, int as_written, b200sp_mat *A) { API_BEGIN use_device(da->d.ctx); *A = wrap(da->d.ctx, assemble_stress(da->d, as_written)); API_END }
int b200sp_assemble_rhs(b200sp_dmda da, int as_written, int rhs_kind, b200sp_vec f) {
  API_BEGIN
  B2_REQUIRE(f->v.n >= (int64_t)da->d.xm * da->d.ym * 2, "assemble_rhs: vector too short");
  assemble_rhs(da->d, as_written, rhs_kind, f->v.d);
  API_END
}
int b200sp_assemble_kkt(b200sp_dmda da, b200sp_mat *Bt, b200sp_mat *B, b200sp_mat *C, b200sp_mat *Q) {
  API_BEGIN
  std::shared_ptr<Csr> bt, b, c, q;
  use_device(da->d.ctx);
  assemble_kkt(da->d, Bt ? &bt : nullptr, B ? &b : nullptr, C ? &c : nullptr, Q ? &q : nullptr);
  if (Bt) *Bt = wrap(da->d.ctx, bt);
  if (B) *B = wrap(da->d.ctx, b);
  if (C) *C = wrap(da->d.ctx, c);
  if (Q) *Q = wrap(da->d.ctx, q);
  API_END
}
int b200sp_assemble_constraints(b200sp_dmda da, b200sp_mat *B, b200sp_mat *Bt) {
  API_BEGIN
  B2_REQUIRE(da && B, "assemble_constraints: bad arguments");
  use_device(da->d.ctx);
  std::shared_ptr<Csr> b, bt;
  assemble_constraints(da->d, &b, Bt ? &bt : nullptr);
  *B = wrap(da->d.ctx, b);
  if (Bt) *Bt = wrap(da->d.ctx, bt);
  API_END
}
int b200sp_interp_q1(b200sp_ctx ctx, int Mc, int Nc, int dof, int bc, b200sp_mat *P) { API_BEGIN *P = wrap(&ctx->c, interp_q1(&ctx->c, Mc, Nc, dof, bc)); API_END }

// ---------------------------------------------------------------- 3-D DMDA + assembly (BASELINE config 4)
// DMDACreate3d(PETSC_DECIDE x3) process grid (PETSc da3.c: squarish factorisation), ownership M/m + (M%m > i) per direction
static void dmda3_proc_grid(int M, int N, int P, int size, int *pm, int *pn, int *pp) {
  int n = (int)(0.5 + std::pow(((double)N * N) * ((double)size) / ((double)P * M), 1.0 / 3.0)), m, p = 1;
  if (!n) n = 1;
  while (n > 0) { const int pmn = size / n; if (n * pmn == size) break; n--; }
  if (!n) n = 1;
  m = (int)(0.5 + std::sqrt(((double)M) * ((double)size) / ((double)P * n)));
  if (!m) m = 1;
  while (m > 0) { p = size / (m * n); if (m * n * p == size) break; m--; }
  if (M > P && m < p) std::swap(m, p);
  *pm = m; *pn = n; *pp = p;
}
int b200sp_dmda3d_proc_grid(int M, int N, int P, int size, int *m, int *n, int *p) {
  API_BEGIN B2_REQUIRE(M > 0 && N > 0 && P > 0 && size > 0, "bad args"); dmda3_proc_grid(M, N, P, size, m, n, p); API_END
}
// host-only index arithmetic of the 3-D partition (no device needed): owned box of a rank, PETSc global id and owner of a node
struct Part3 {
  int m, n, p;
  std::vector<int> lx, ly, lz, xo, yo, zo, rstart;
  Part3(int M, int N, int P, int size) {
    dmda3_proc_grid(M, N, P, size, &m, &n, &p);
    B2_REQUIRE(m * n * p == size && m <= M && n <= N && p <= P, "dmda3d: size does not factor into a process grid for this mesh");
    lx.resize((size_t)m); ly.resize((size_t)n); lz.resize((size_t)p);
    dmda_ownership(M, m, lx.data()); dmda_ownership(N, n, ly.data()); dmda_ownership(P, p, lz.data());
    xo.assign((size_t)m + 1, 0); yo.assign((size_t)n + 1, 0); zo.assign((size_t)p + 1, 0);
    for (int i = 0; i < m; ++i) xo[(size_t)i + 1] = xo[(size_t)i] + lx[(size_t)i];
    for (int i = 0; i < n; ++i) yo[(size_t)i + 1] = yo[(size_t)i] + ly[(size_t)i];
    for (int i = 0; i < p; ++i) zo[(size_t)i + 1] = zo[(size_t)i] + lz[(size_t)i];
    rstart.assign((size_t)size + 1, 0);
    for (int r = 0; r < size; ++r) rstart[(size_t)r + 1] = rstart[(size_t)r] + lx[(size_t)(r % m)] * ly[(size_t)((r / m) % n)] * lz[(size_t)(r / (m * n))];
  }
  void box(int r, int *xs, int *ys, int *zs, int *xm, int *ym, int *zm) const {
    const int pi = r % m, pj = (r / m) % n, pk = r / (m * n);
    *xs = xo[(size_t)pi]; *ys = yo[(size_t)pj]; *zs = zo[(size_t)pk]; *xm = lx[(size_t)pi]; *ym = ly[(size_t)pj]; *zm = lz[(size_t)pk];
  }
  static int owner_of(const std::vector<int> &off, int v) { return (int)(std::upper_bound(off.begin(), off.end(), v) - off.begin()) - 1; }
  int owner(int i, int j, int k) const { return (owner_of(zo, k) * n + owner_of(yo, j)) * m + owner_of(xo, i); }
  int gnode(int i, int j, int k) const {
    const int pi = owner_of(xo, i), pj = owner_of(yo, j), pk = owner_of(zo, k), r = (pk * n + pj) * m + pi;
    return rstart[(size_t)r] + ((k - zo[(size_t)pk]) * ly[(size_t)pj] + (j - yo[(size_t)pj])) * lx[(size_t)pi] + (i - xo[(size_t)pi]);
  }
};
int b200sp_dmda3d_corners(int M, int N, int P, int size, int rank, int *xs, int *ys, int *zs, int *xm, int *ym, int *zm) {
  API_BEGIN
  B2_REQUIRE(M > 0 && N > 0 && P > 0 && size > 0 && rank >= 0 && rank < size, "dmda3d_corners: bad arguments");
  Part3(M, N, P, size).box(rank, xs, ys, zs, xm, ym, zm);
  API_END
}
int b200sp_dmda3d_global_node(int M, int N, int P, int size, int i, int j, int k, int *gnode, int *owner) {
  API_BEGIN
  B2_REQUIRE(i >= 0 && i < M && j >= 0 && j < N && k >= 0 && k < P && size > 0, "dmda3d_global_node: node out of range");
  Part3 pt(M, N, P, size);
  if (gnode) *gnode = pt.gnode(i, j, k);
  if (owner) *owner = pt.owner(i, j, k);
  API_END
}
int b200sp_dmda3d_create(b200sp_ctx ctx, int M, int N, int P, b200sp_dmda3d *da) {
  API_BEGIN
  B2_REQUIRE(ctx && da && M >= 2 && N >= 2 && P >= 2, "dmda3d_create: bad arguments");
  Ctx *c = &ctx->c;
  use_device(c);
  auto *h = new b200sp_dmda3d_s();
  try {
    Dmda3 &d = h->d;
    d.ctx = c; d.M = M; d.N = N; d.P = P;
    const int size = c->size, rank = c->rank;
    const Part3 part(M, N, P, size);
    d.pm = part.m; d.pn = part.n; d.pp = part.p;
    const std::vector<int> &rstart = part.rstart;
    part.box(rank, &d.xs, &d.ys, &d.zs, &d.xm, &d.ym, &d.zm);
    B2_REQUIRE(size == 1 || (d.xm >= 2 && d.ym >= 2 && d.zm >= 2), "dmda3d: every rank must own at least 2 x 2 x 2 nodes");
    d.g0 = rstart[(size_t)rank];
    auto gnode = [&](int i, int j, int k) { return part.gnode(i, j, k); };
    // ghost nodes: the one-node layer around the owned box, clipped to the domain, sorted by global id (MPIAIJ garray order)
    const int ex = d.xm + 2, ey = d.ym + 2, ez = d.zm + 2;
    std::vector<std::pair<int, int>> gh; // (global node id, ext index)
    std::vector<int> lut((size_t)ex * ey * ez, -1);
    for (int k = d.zs - 1; k <= d.zs + d.zm; ++k)
      for (int j = d.ys - 1; j <= d.ys + d.ym; ++j)
        for (int i = d.xs - 1; i <= d.xs + d.xm; ++i) {
          if (i < 0 || i >= M || j < 0 || j >= N || k < 0 || k >= P) continue;
          const int e = ((k - d.zs + 1) * ey + (j - d.ys + 1)) * ex + (i - d.xs + 1);
          const bool owned = i >= d.xs && i < d.xs + d.xm && j >= d.ys && j < d.ys + d.ym && k >= d.zs && k < d.zs + d.zm;
          if (owned) lut[(size_t)e] = ((k - d.zs) * d.ym + (j - d.ys)) * d.xm + (i - d.xs);
          else gh.push_back({gnode(i, j, k), e});
        }
    std::sort(gh.begin(), gh.end());
    const int nown = d.xm * d.ym * d.zm;
    std::vector<int> ghost_gnode;
    for (size_t t = 0; t < gh.size(); ++t) { lut[(size_t)gh[t].second] = nown + (int)t; ghost_gnode.push_back(gh[t].first); }
    dmda3_build_lut(d, lut);
    if (size > 1) {
      d.layout = std::make_shared<Layout>();
      d.layout->M = M; d.layout->N = N; d.layout->size = size; d.layout->rstart = rstart;
      d.halo = make_halo_general(c, *d.layout, rank, ghost_gnode);
    }
  } catch (...) { delete h; throw; }
  *da = h;
  API_END
}
int b200sp_dmda3d_destroy(b200sp_dmda3d da) { API_BEGIN delete da; API_END }
int b200sp_dmda3d_get_info(b200sp_dmda3d da, int *xs, int *ys, int *zs, int *xm, int *ym, int *zm, int64_t *gstart) {
  API_BEGIN
  const Dmda3 &d = da->d;
  if (xs) *xs = d.xs; if (ys) *ys = d.ys; if (zs) *zs = d.zs; if (xm) *xm = d.xm; if (ym) *ym = d.ym; if (zm) *zm = d.zm;
  if (gstart) *gstart = d.g0;
  API_END
}
int b200sp_dmda3d_bc_ids(b200sp_dmda3d da, int dof, int *n, int *ids) {
  API_BEGIN
  std::vector<int> v = dmda3_bc_ids(da->d, dof);
  if (n) *n = (int)v.size();
  if (ids) std::copy(v.begin(), v.end(), ids);
  API_END
}
int b200sp_assemble3d_stress(b200sp_dmda3d da, b200sp_mat *A) { API_BEGIN use_device(da->d.ctx); *A = wrap(da->d.ctx, assemble3_stress(da->d)); API_END }
int b200sp_assemble3d_rhs(b200sp_dmda3d da, int rhs_kind, b200sp_vec f) {
  API_BEGIN
  B2_REQUIRE(f->v.n >= (int64_t)da->d.xm * da->d.ym * da->d.zm * 3, "assemble3d_rhs: vector too short");
  use_device(da->d.ctx);
  assemble3_rhs(da->d, rhs_kind, f->v.d);
  API_END
}
int b200sp_assemble3d_kkt(b200sp_dmda3d da, b200sp_mat *Bt, b200sp_mat *B, b200sp_mat *C, b200sp_mat *Q) {
  API_BEGIN
  std::shared_ptr<Csr> bt, b, c, q;
  use_device(da->d.ctx);
  assemble3_kkt(da->d, Bt ? &bt : nullptr, B ? &b : nullptr, C ? &c : nullptr, Q ? &q : nullptr);
  if (Bt) *Bt = wrap(da->d.ctx, bt);
  if (B) *B = wrap(da->d.ctx, b);
  if (C) *C = wrap(da->d.ctx, c);
  if (Q) *Q = wrap(da->d.ctx, q);
  API_END
}

// ---------------------------------------------------------------- KSP
int b200sp_ksp_create(b200sp_ctx ctx, b200sp_ksp *ksp) { API_BEGIN B2_REQUIRE(ctx && ksp, "ksp_create: bad arguments"); *ksp = new b200sp_ksp_s(&ctx->c); API_END }
int b200sp_ksp_destroy(b200sp_ksp *ksp) {
  API_BEGIN
  if (ksp && *ksp) { (*ksp)->s.ctx->sync(); delete *ksp; *ksp = nullptr; }
  API_END
}
int b200sp_ksp_set_operators(b200sp_ksp ksp, b200sp_mat Amat, b200sp_mat Pmat) {
  API_BEGIN
  B2_REQUIRE(Amat && Pmat, "KSPSetOperators: null matrix");
  use_device(ksp->s.ctx); ksp->s.set_operators(Amat->m, Pmat->m);
  API_END
}
int b200sp_ksp_set_options(b200sp_ksp ksp, const char *options) { API_BEGIN ksp->s.set_options(options); API_END }
int b200sp_ksp_set_schur_user_mat(b200sp_ksp ksp, b200sp_mat Q) { API_BEGIN ksp->s.schur_user = plain(Q).ctx ? Q->m.csr : nullptr; ksp->s.is_setup = false; API_END }
int b200sp_ksp_set_dmda(b200sp_ksp ksp, b200sp_dmda da) { API_BEGIN ksp->s.have_grid = true; ksp->s.grid_M = da->d.M; ksp->s.grid_N = da->d.N; API_END }
int b200sp_ksp_setup(b200sp_ksp ksp) { API_BEGIN use_device(ksp->s.ctx); ksp->s.setup(); API_END }
int b200sp_ksp_solve(b200sp_ksp ksp, b200sp_vec b, b200sp_vec x) {
  API_BEGIN
  use_device(ksp->s.ctx);
  if (!ksp->s.current()) ksp->s.setup(); // also when matrix values changed since KSPSetUp (PETSc: object state)
  B2_REQUIRE(b->v.n == ksp->s.outer->n && x->v.n == b->v.n && b != x, "KSPSolve: size mismatch or aliasing");
  ksp->s.outer->solve(b->v.d, x->v.d, false);
  ksp->s.ctx->sync();
  check_device_error(ksp->s.ctx);
  API_END
}
int b200sp_ksp_solve_host(b200sp_ksp ksp, const double *b_host, double *x_host, int64_t n) {
  API_BEGIN
  use_device(ksp->s.ctx);
  if (!ksp->s.current()) ksp->s.setup(); // also when matrix values changed since KSPSetUp (PETSc: object state)
  Ctx *c = ksp->s.ctx;
  B2_REQUIRE(n == ksp->s.outer->n, "KSPSolve(host): size mismatch");
  // staging vectors live with the KSP (allocated once): the call itself only copies and solves
  DevBuf<double> &b = ksp->s.host_b, &x = ksp->s.host_x;
  if (b.n < (size_t)n + 2) { b.alloc((size_t)n + 2); x.alloc((size_t)n + 2); }
  B2_CUDA(cudaMemcpyAsync(b.p, b_host, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  ksp->s.outer->solve(b.p, x.p, false);
  B2_CUDA(cudaMemcpyAsync(x_host, x.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  c->sync();
  check_device_error(c);
  API_END
}
int b200sp_ksp_get_iteration_number(b200sp_ksp ksp, int *its) { API_BEGIN B2_REQUIRE(ksp->s.outer, "KSP not set up"); *its = ksp->s.outer->its; API_END }
int b200sp_ksp_get_residual_norm(b200sp_ksp ksp, double *rnorm) { API_BEGIN B2_REQUIRE(ksp->s.outer, "KSP not set up"); *rnorm = ksp->s.outer->rnorm; API_END }
int b200sp_ksp_get_converged_reason(b200sp_ksp ksp, int *reason) { API_BEGIN B2_REQUIRE(ksp->s.outer, "KSP not set up"); *reason = ksp->s.outer->reason; API_END }
int b200sp_ksp_get_residual_history(b200sp_ksp ksp, double *hist, int cap, int *len) {
  API_BEGIN
  B2_REQUIRE(ksp->s.outer, "KSP not set up");
  const auto &h = ksp->s.outer->hist;
  if (len) *len = (int)h.size();
  if (hist) for (int i = 0; i < cap && i < (int)h.size(); ++i) hist[i] = h[(size_t)i];
  API_END
}
int b200sp_ksp_pc_apply(b200sp_ksp ksp, b200sp_vec x, b200sp_vec y) {
  API_BEGIN
  if (!ksp->s.current()) ksp->s.setup(); // also when matrix values changed since KSPSetUp (PETSc: object state)
  B2_REQUIRE(x->v.n == ksp->s.outer->n && y->v.n == x->v.n && x != y, "PCApply: size mismatch or aliasing");
  if (ksp->s.outer_pc) ksp->s.outer_pc->apply(x->v.d, y->v.d);
  else vec_copy(ksp->s.ctx, x->v.n, x->v.d, y->v.d);
  ksp->s.ctx->sync();
  check_device_error(ksp->s.ctx);
  API_END
}
// ---- PC as an object of its own (PCCreate / PCSetOperators / PCSetFromOptions / PCSetUp / PCApply / PCDestroy)
int b200sp_pc_create(b200sp_ctx ctx, b200sp_pc *pc) { API_BEGIN B2_REQUIRE(ctx && pc, "pc_create: bad arguments"); *pc = new b200sp_pc_s(&ctx->c); API_END }
int b200sp_pc_destroy(b200sp_pc *pc) {
  API_BEGIN
  if (pc && *pc) { (*pc)->s.ctx->sync(); delete *pc; *pc = nullptr; }
  API_END
}
int b200sp_pc_set_operators(b200sp_pc pc, b200sp_mat Amat, b200sp_mat Pmat) {
  API_BEGIN
  B2_REQUIRE(pc && Amat && Pmat, "PCSetOperators: null argument");
  use_device(pc->s.ctx); pc->s.set_operators(Amat->m, Pmat->m);
  API_END
}
int b200sp_pc_set_options(b200sp_pc pc, const char *options) { API_BEGIN pc->s.set_options(options); pc->s.is_setup = false; API_END }
int b200sp_pc_set_schur_user_mat(b200sp_pc pc, b200sp_mat Q) { API_BEGIN pc->s.schur_user = plain(Q).ctx ? Q->m.csr : nullptr; pc->s.is_setup = false; API_END }
int b200sp_pc_set_dmda(b200sp_pc pc, b200sp_dmda da) { API_BEGIN pc->s.have_grid = true; pc->s.grid_M = da->d.M; pc->s.grid_N = da->d.N; API_END }
int b200sp_pc_setup(b200sp_pc pc) { API_BEGIN use_device(pc->s.ctx); pc->s.setup(); API_END }
int b200sp_pc_apply(b200sp_pc pc, b200sp_vec x, b200sp_vec y) {
  API_BEGIN
  if (!pc->s.current()) pc->s.setup();
  B2_REQUIRE(x->v.n == pc->s.outer->n && y->v.n == x->v.n && x != y, "PCApply: size mismatch or aliasing");
  if (pc->s.outer_pc) pc->s.outer_pc->apply(x->v.d, y->v.d);
  else vec_copy(pc->s.ctx, x->v.n, x->v.d, y->v.d); // PCNONE
  pc->s.ctx->sync();
  check_device_error(pc->s.ctx);
  API_END
}
int b200sp_pc_view(b200sp_pc pc, char *buf, int buflen) {
  API_BEGIN
  B2_REQUIRE(buf && buflen > 0, "pc_view: no buffer");
  const std::string s = pc->s.outer_pc ? pc->s.outer_pc->view(0) : std::string(pc->s.is_setup ? "PC none\n" : "PC not set up\n");
  const size_t n = std::min(s.size(), (size_t)buflen - 1);
  std::memcpy(buf, s.c_str(), n);
  buf[n] = 0;
  API_END
}
int b200sp_ksp_view(b200sp_ksp ksp, char *buf, int buflen) {
  API_BEGIN
  std::string s = ksp->s.view();
  B2_REQUIRE(buf && buflen > 0, "ksp_view: no buffer");
  const size_t n = std::min(s.size(), (size_t)buflen - 1);
  std::memcpy(buf, s.c_str(), n);
  buf[n] = 0;
  API_END
}

} // extern "C"
