// Aggregation multigrid set-up on the device (-pc_type gamg; SURVEY.md section 8(f) rank 1: AMG V-cycle for A00 and
// for L = B B^T in LSC, PETSc analogue PCGAMG -pc_gamg_type agg).  The algorithm is the one oracle/sp_oracle_amg.c
// defines -- every step is a Jacobi-style pass (reads the previous pass only), so the result does not depend on the
// schedule and the aggregates / tentative prolongator are bit-identical to the oracle's:
//   strength graph on nodes -> distance-2 maximal independent set by hashed priority (Bell, Dalton, Olson 2012)
//   -> root + neighbours, then the rest joins its highest-key aggregated neighbour -> P_t = sqrt(w_i / W_aggregate)
//   -> P = P_t - omega D^-1 A P_t (SpGEMM) -> A_c = P^T A P (SpGEMM).
// The Dirichlet rows the reference leaves as identity rows (src/Discretization.c:268, MatZeroRowsColumns) have no
// neighbours and stay out of the coarse space; the level smoother handles them.
#include "core.h"
#include <cstdlib>

namespace b200sp {
namespace {

// priority of a node, distinct per node, < 2^62.  order 0: hashed (pseudo-random permutation, ~10 rounds); order 1: the
// node number (greedy in descending natural order: a regular root lattice on lexicographically numbered grids, at
// the price of O(grid side) rounds)
__host__ __device__ inline unsigned long long amg_key(int i, int order) {
  if (order == 1) return (unsigned long long)(unsigned int)i + 1ull;
  unsigned int h = (unsigned int)i * 2654435761u;
  h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13;
  return ((unsigned long long)(h >> 2) << 32) | (unsigned int)i;
}
constexpr unsigned long long ST_UNDECIDED = 1ull << 62, ST_ROOT = 2ull << 62;

inline int amg_grid(Ctx *c, int64_t n) {
  int64_t g = (n + 255) / 256;
  const int64_t cap = (int64_t)c->num_sms * 16;
  return (int)std::max<int64_t>(1, std::min(g, cap));
}

// one merged pass over the bs (<= 4) sorted rows of node i: f(j, s) once per block column j, s = sum |a| row-major
template <class F>
__device__ inline void walk_node(const int *__restrict__ rp, const int *__restrict__ col, const double *__restrict__ val, int bs, int i, F f) {
  int p[4], e[4];
  for (int c = 0; c < 4; ++c) { p[c] = 0; e[c] = 0; }
  for (int c = 0; c < bs; ++c) { p[c] = rp[i * bs + c]; e[c] = rp[i * bs + c + 1]; }
  for (;;) {
    int j = -1;
    for (int c = 0; c < bs; ++c)
      if (p[c] < e[c]) { const int n = col[p[c]] / bs; if (j < 0 || n < j) j = n; }
    if (j < 0) break;
    double s = 0.0;
    for (int c = 0; c < bs; ++c)
      while (p[c] < e[c] && col[p[c]] / bs == j) { s += fabs(val[p[c]]); ++p[c]; }
    f(j, s);
  }
}

__global__ void __launch_bounds__(256) k_amg_diag(int nn, int bs, const int *__restrict__ rp, const int *__restrict__ col, const double *__restrict__ val, double *sd) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += gridDim.x * blockDim.x) {
    double d = 0.0;
    walk_node(rp, col, val, bs, i, [&](int j, double s) { if (j == i) d = s; });
    sd[i] = d;
  }
}
// FILL = false: cnt[i] = number of neighbours ; FILL = true: gcol[grp[i] ..] = neighbours (ascending)
template <bool FILL>
__global__ void __launch_bounds__(256) k_amg_graph(int nn, int bs, const int *__restrict__ rp, const int *__restrict__ col, const double *__restrict__ val,
                                                   const double *__restrict__ sd, double theta, const int *__restrict__ grp, int *gcol, int *cnt) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += gridDim.x * blockDim.x) {
    int n = 0;
    const int base = FILL ? grp[i] : 0;
    const double sdi = sd[i];
    walk_node(rp, col, val, bs, i, [&](int j, double s) {
      if (j == i) return;
      const bool keep = theta > 0.0 ? (s > theta * sqrt(sdi * sd[j])) : (s > 0.0);
      if (keep) { if (FILL) gcol[base + n] = j; ++n; }
    });
    if (!FILL) cnt[i] = n;
  }
}
__global__ void __launch_bounds__(256) k_amg_init(int nn, int order, const int *__restrict__ grp, unsigned long long *t, int *counter) {
  for (int base = blockIdx.x * blockDim.x; base < nn; base += gridDim.x * blockDim.x) {
    const int i = base + threadIdx.x;
    bool u = false;
    if (i < nn) { u = grp[i + 1] > grp[i]; t[i] = u ? (ST_UNDECIDED | amg_key(i, order)) : 0ull; }
    const int c = __syncthreads_count(u);
    if (threadIdx.x == 0 && c) atomicAdd(counter, c);
  }
}
__global__ void __launch_bounds__(256) k_amg_prop(int nn, const int *__restrict__ grp, const int *__restrict__ gcol, const unsigned long long *__restrict__ in, unsigned long long *out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += gridDim.x * blockDim.x) {
    unsigned long long m = in[i];
    for (int k = grp[i]; k < grp[i + 1]; ++k) { const unsigned long long v = in[gcol[k]]; if (v > m) m = v; }
    out[i] = m;
  }
}
// undecided node: largest key among the undecided within distance 2 and no root there -> root; a root there -> out
__global__ void __launch_bounds__(256) k_amg_decide(int nn, int order, unsigned long long *t, const unsigned long long *__restrict__ m2, int *counter) {
  for (int base = blockIdx.x * blockDim.x; base < nn; base += gridDim.x * blockDim.x) {
    const int i = base + threadIdx.x;
    bool u = false;
    if (i < nn) {
      const unsigned long long ti = t[i];
      if ((ti >> 62) == 1ull) {
        const unsigned long long m = m2[i];
        if (m == ti) t[i] = ST_ROOT | amg_key(i, order);
        else if ((m >> 62) == 2ull) t[i] = 0ull;
        else u = true;
      }
    }
    const int c = __syncthreads_count(u);
    if (threadIdx.x == 0 && c) atomicAdd(counter, c);
  }
}
__global__ void __launch_bounds__(256) k_amg_rootflag(int nn, const unsigned long long *__restrict__ t, int *flag) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += gridDim.x * blockDim.x) flag[i] = (t[i] >> 62) == 2ull ? 1 : 0;
}
__global__ void __launch_bounds__(256) k_amg_join1(int nn, int order, const int *__restrict__ grp, const int *__restrict__ gcol, const unsigned long long *__restrict__ t,
                                                   const int *__restrict__ rootid, int *agg1) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += gridDim.x * blockDim.x) {
    int a = -1;
    if ((t[i] >> 62) == 2ull) a = rootid[i];
    else {
      unsigned long long best = 0ull;
      int who = -1;
      for (int k = grp[i]; k < grp[i + 1]; ++k) {
        const int j = gcol[k];
        if ((t[j] >> 62) == 2ull && amg_key(j, order) >= best) { best = amg_key(j, order); who = j; }
      }
      if (who >= 0) a = rootid[who];
    }
    agg1[i] = a;
  }
}
__global__ void __launch_bounds__(256) k_amg_join2(int nn, int order, const int *__restrict__ grp, const int *__restrict__ gcol, const int *__restrict__ agg1, int *agg) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += gridDim.x * blockDim.x) {
    int a = agg1[i];
    if (a < 0) {
      unsigned long long best = 0ull;
      int who = -1;
      for (int k = grp[i]; k < grp[i + 1]; ++k) {
        const int j = gcol[k];
        if (agg1[j] >= 0 && amg_key(j, order) >= best) { best = amg_key(j, order); who = j; }
      }
      if (who >= 0) a = agg1[who];
    }
    agg[i] = a;
  }
}
// W[a] = sum of the node weights of aggregate a (integers: the order of the atomics does not matter)
__global__ void __launch_bounds__(256) k_amg_sizes(int nn, const int *__restrict__ agg, const int *__restrict__ w, int *W) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += gridDim.x * blockDim.x)
    if (agg[i] >= 0) atomicAdd(&W[agg[i]], w ? w[i] : 1);
}
__global__ void __launch_bounds__(256) k_amg_rowcnt(int nn, int bs, const int *__restrict__ agg, int *cnt) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nn * bs; r += gridDim.x * blockDim.x) cnt[r] = agg[r / bs] >= 0 ? 1 : 0;
}
__global__ void __launch_bounds__(256) k_amg_tentative(int nn, int bs, const int *__restrict__ agg, const int *__restrict__ w, const int *__restrict__ W,
                                                       const int *__restrict__ rp, int *col, double *val) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nn * bs; r += gridDim.x * blockDim.x) {
    const int a = agg[r / bs];
    if (a >= 0) { col[rp[r]] = a * bs + r % bs; val[rp[r]] = sqrt((double)(w ? w[r / bs] : 1) / (double)W[a]); }
  }
}
__global__ void __launch_bounds__(256) k_scale_rows(int nrows, const int *__restrict__ rp, const double *__restrict__ val, const double *__restrict__ d, double *out) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x) {
    const double s = d[r];
    for (int k = rp[r]; k < rp[r + 1]; ++k) out[k] = s * val[k];
  }
}
} // namespace

int amg_aggregate(const Csr &A, int bs, double theta, int order, DevBuf<int> &agg) {
  Ctx *c = A.ctx;
  B2_REQUIRE(!A.halo, "gamg: row-partitioned matrices are not aggregated (single-rank set-up only)");
  B2_REQUIRE(order == 0 || order == 1, "gamg: unknown independent-set ordering");
  B2_REQUIRE(bs >= 1 && bs <= 4 && A.nrows == A.ncols && A.nrows % bs == 0, "gamg: square matrix with block size 1..4 expected");
  const int nn = A.nrows / bs;
  agg.alloc((size_t)nn + 1);
  if (nn == 0) return 0;
  const int g = amg_grid(c, nn);
  DevBuf<double> sd((size_t)nn);
  DevBuf<int> cnt((size_t)nn + 1), grp((size_t)nn + 1);
  int gnnz = 0;
  {
    LaunchScope ls(c, "setup");
    k_amg_diag<<<g, 256, 0, c->stream>>>(nn, bs, A.rowptr.p, A.col.p, A.val.p, sd.p);
    k_amg_graph<false><<<g, 256, 0, c->stream>>>(nn, bs, A.rowptr.p, A.col.p, A.val.p, sd.p, theta, nullptr, nullptr, cnt.p);
    check_launch("k_amg_graph");
  }
  exclusive_scan_i32(c, cnt.p, grp.p, nn, &gnnz);
  DevBuf<int> gcol((size_t)gnnz + 1);
  DevBuf<unsigned long long> t((size_t)nn), m1((size_t)nn), m2((size_t)nn);
  DevBuf<int> counter(1);
  int undecided = 0;
  {
    LaunchScope ls(c, "setup");
    k_amg_graph<true><<<g, 256, 0, c->stream>>>(nn, bs, A.rowptr.p, A.col.p, A.val.p, sd.p, theta, grp.p, gcol.p, nullptr);
    B2_CUDA(cudaMemsetAsync(counter.p, 0, sizeof(int), c->stream));
    k_amg_init<<<g, 256, 0, c->stream>>>(nn, order, grp.p, t.p, counter.p);
    check_launch("k_amg_init");
  }
  B2_CUDA(cudaMemcpyAsync(&undecided, counter.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  c->sync();
  // rounds per host read-back of the undecided count: a round after the last decision changes nothing, so running a
  // few too many is harmless and saves a synchronisation per round (the natural ordering needs O(grid side) rounds)
  const char *eb = getenv("B200SP_AMG_BATCH"); // rounds per read-back (tuning knob)
  const int batch = eb && atoi(eb) > 0 ? atoi(eb) : (order == 1 ? 8 : 2);
  for (int round = 0; undecided > 0; round += batch) {
    B2_REQUIRE(round < 200000, "gamg: the independent-set selection did not terminate");
    LaunchScope ls(c, "setup");
    for (int b = 0; b < batch; ++b) {
      B2_CUDA(cudaMemsetAsync(counter.p, 0, sizeof(int), c->stream));
      k_amg_prop<<<g, 256, 0, c->stream>>>(nn, grp.p, gcol.p, t.p, m1.p);
      k_amg_prop<<<g, 256, 0, c->stream>>>(nn, grp.p, gcol.p, m1.p, m2.p);
      k_amg_decide<<<g, 256, 0, c->stream>>>(nn, order, t.p, m2.p, counter.p);
    }
    check_launch("k_amg_decide");
    B2_CUDA(cudaMemcpyAsync(&undecided, counter.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    c->sync();
  }
  DevBuf<int> flag((size_t)nn + 1), rootid((size_t)nn + 1), agg1((size_t)nn + 1);
  int nagg = 0;
  {
    LaunchScope ls(c, "setup");
    k_amg_rootflag<<<g, 256, 0, c->stream>>>(nn, t.p, flag.p);
    check_launch("k_amg_rootflag");
  }
  exclusive_scan_i32(c, flag.p, rootid.p, nn, &nagg);
  {
    LaunchScope ls(c, "setup");
    k_amg_join1<<<g, 256, 0, c->stream>>>(nn, order, grp.p, gcol.p, t.p, rootid.p, agg1.p);
    k_amg_join2<<<g, 256, 0, c->stream>>>(nn, order, grp.p, gcol.p, agg1.p, agg.p);
    check_launch("k_amg_join");
  }
  c->sync();
  return nagg;
}

// w: finest-level nodes behind every node of this level (null: ones) -- the finest level's constant vector is sqrt(w)
// here, so P_t[(i,c),(a,c)] = sqrt(w_i / W_a); W (the coarse level's weights) is returned in wc
std::shared_ptr<Csr> amg_tentative(Ctx *c, int nn, int bs, const DevBuf<int> &agg, int nagg, const int *w, DevBuf<int> &wc) {
  const int nrows = nn * bs;
  DevBuf<int> cnt((size_t)nrows + 1), rp((size_t)nrows + 1);
  DevBuf<int> &size = wc;
  size.alloc((size_t)nagg + 1);
  int nnz = 0;
  B2_CUDA(cudaMemsetAsync(size.p, 0, sizeof(int) * ((size_t)nagg + 1), c->stream));
  if (nrows) {
    LaunchScope ls(c, "setup");
    k_amg_sizes<<<amg_grid(c, nn), 256, 0, c->stream>>>(nn, agg.p, w, size.p);
    k_amg_rowcnt<<<amg_grid(c, nrows), 256, 0, c->stream>>>(nn, bs, agg.p, cnt.p);
    check_launch("k_amg_rowcnt");
  }
  exclusive_scan_i32(c, cnt.p, rp.p, nrows, &nnz);
  auto P = csr_alloc_public(c, nrows, nagg * bs, nnz);
  B2_CUDA(cudaMemcpyAsync(P->rowptr.p, rp.p, sizeof(int) * ((size_t)nrows + 1), cudaMemcpyDeviceToDevice, c->stream));
  if (nrows) {
    LaunchScope ls(c, "setup");
    k_amg_tentative<<<amg_grid(c, nrows), 256, 0, c->stream>>>(nn, bs, agg.p, w, size.p, P->rowptr.p, P->col.p, P->val.p);
    check_launch("k_amg_tentative");
  }
  c->sync();
  P->plan();
  return P;
}

std::shared_ptr<Csr> csr_scale_rows(const Csr &A, const double *d) { // diag(d) * A, owned rows (valid for row-partitioned A too)
  Ctx *c = A.ctx;
  auto C = csr_alloc_public(c, A.nrows, A.ncols, A.nnz);
  B2_CUDA(cudaMemcpyAsync(C->rowptr.p, A.rowptr.p, sizeof(int) * ((size_t)A.nrows + 1), cudaMemcpyDeviceToDevice, c->stream));
  if (A.nnz) B2_CUDA(cudaMemcpyAsync(C->col.p, A.col.p, sizeof(int) * (size_t)A.nnz, cudaMemcpyDeviceToDevice, c->stream));
  if (A.nrows) {
    LaunchScope ls(c, "setup");
    k_scale_rows<<<amg_grid(c, A.nrows), 256, 0, c->stream>>>(A.nrows, A.rowptr.p, A.val.p, d, C->val.p);
    check_launch("k_scale_rows");
  }
  c->sync();
  C->grid_M = A.grid_M; C->grid_N = A.grid_N; C->dof_r = A.dof_r; C->dof_c = A.dof_c;
  if (A.halo) csr_copy_distribution(*C, A);
  C->plan();
  return C;
}

// P = P_t - omega D^-1 (A P_t), D = diag(A) with 0 -> 1 (PCJACOBI's rule)
std::shared_ptr<Csr> amg_smooth_prolongator(const Csr &A, const Csr &Pt, double omega) {
  Ctx *c = A.ctx;
  DevBuf<double> dinv((size_t)A.nrows + 1);
  csr_get_diagonal(A, dinv.p);
  vec_reciprocal_safe(c, A.nrows, dinv.p);
  auto AP = csr_matmat(A, Pt);
  auto DAP = csr_scale_rows(*AP, dinv.p);
  return csr_add_scaled(Pt, -omega, *DAP);
}
} // namespace b200sp
