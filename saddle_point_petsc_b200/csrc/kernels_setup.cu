// kernels_setup.cu -- one-off setup kernels: prefix scan, CSR construction from host CSR / COO (stable
// radix sort + in-order duplicate sum = MatSetValues(ADD_VALUES)+MatAssemblyEnd, src/Discretization.c:165-169),
// explicit transpose, SpGEMM (MatMatMult used by selfp / LSC), small dense inverse
// (coarsest multigrid level).
#include "dev.cuh"
#include "dist.h"
#include <algorithm>

namespace b200sp {

namespace {

// ---------------------------------------------------------------- exclusive scan (int32)
constexpr int SCAN_ITEMS = 4;                    // per thread
constexpr int SCAN_TILE = 256 * SCAN_ITEMS;      // per CTA

__global__ void __launch_bounds__(256) k_scan_tiles(const int *__restrict__ in, int *__restrict__ out, int64_t n, int *tile_sums) {
  __shared__ int s_warp[8];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS], t = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    v[k] = base + k < n ? in[base + k] : 0;
    t += v[k];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = t;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int u = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += u;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  int woff = 0;
  for (int w = 0; w < warp; ++w) woff += s_warp[w];
  int excl = woff + incl - t;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    if (base + k < n) out[base + k] = excl;
    excl += v[k];
  }
  if (threadIdx.x == 255) tile_sums[blockIdx.x] = woff + incl;
}
__global__ void __launch_bounds__(256) k_scan_add(int *__restrict__ out, int64_t n, const int *__restrict__ tile_offs) {
  const int off = tile_offs[blockIdx.x];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k)
    if (base + k < n) out[base + k] += off;
}

void scan_rec(Ctx *c, const int *in, int *out, int64_t n) { // out[i] = sum_{j<i} in[j], i < n
  if (n <= 0) return;
  int64_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  DevBuf<int> sums((size_t)ntiles), offs((size_t)ntiles);
  {
    LaunchScope ls(c, "setup");
    k_scan_tiles<<<(unsigned)ntiles, 256, 0, c->stream>>>(in, out, n, sums.p);
    check_launch("k_scan_tiles");
  }
  if (ntiles > 1) {
    scan_rec(c, sums.p, offs.p, ntiles);
    LaunchScope ls(c, "setup");
    k_scan_add<<<(unsigned)ntiles, 256, 0, c->stream>>>(out, n, offs.p);
    check_launch("k_scan_add");
  }
  c->sync(); // temporaries die here
}

__global__ void k_last_total(const int *in, const int *out, int64_t n, int *total) { *total = out[n - 1] + in[n - 1]; }

} // namespace

// out has n+1 entries; out[n] = total.  `in` and `out` must not alias.
void exclusive_scan_i32(Ctx *c, const int *in, int *out, int64_t n, int *total_host) {
  int total = 0;
  if (n > 0) {
    scan_rec(c, in, out, n);
    LaunchScope ls(c, "setup");
    k_last_total<<<1, 1, 0, c->stream>>>(in, out, n, out + n);
    check_launch("k_last_total");
    B2_CUDA(cudaMemcpyAsync(&total, out + n, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    c->sync();
  } else {
    B2_CUDA(cudaMemsetAsync(out, 0, sizeof(int), c->stream));
  }
  if (total_host) *total_host = total;
}

static std::shared_ptr<Csr> csr_alloc(Ctx *c, int nrows, int ncols, int64_t nnz) {
  // row pointers and entry offsets are 32-bit (PetscInt in the reference's build): one rank holds < 2^31 entries;
  // larger problems are row-partitioned (per-rank local indexing)
  if (nnz < 0 || nnz >= (int64_t)2147483647 - CSR_PAD)
    throw Error(B200SP_ERR_UNSUPPORTED, "matrix with " + std::to_string(nnz) + " stored entries on one rank: 32-bit row pointers hold < 2^31; use more ranks");
  auto A = std::make_shared<Csr>();
  A->ctx = c;
  A->nrows = nrows;
  A->ncols = ncols;
  A->nnz = nnz;
  A->rowptr.alloc((size_t)nrows + 1);
  A->col.alloc((size_t)nnz + CSR_PAD);
  A->val.alloc((size_t)nnz + CSR_PAD);
  B2_CUDA(cudaMemsetAsync(A->col.p + nnz, 0, sizeof(int) * CSR_PAD, c->stream));
  B2_CUDA(cudaMemsetAsync(A->val.p + nnz, 0, sizeof(double) * CSR_PAD, c->stream));
  return A;
}
std::shared_ptr<Csr> csr_alloc_public(Ctx *c, int nrows, int ncols, int64_t nnz) { return csr_alloc(c, nrows, ncols, nnz); }

std::shared_ptr<Csr> csr_from_host(Ctx *c, int nrows, int ncols, const int *rowptr, const int *col, const double *val) {
  B2_REQUIRE(nrows >= 0 && ncols >= 0 && rowptr, "csr_from_host: bad arguments");
  int64_t nnz = rowptr[nrows];
  for (int r = 0; r < nrows; ++r) {
    B2_REQUIRE(rowptr[r + 1] >= rowptr[r], "csr_from_host: rowptr not monotone");
    for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) {
      B2_REQUIRE(col[k] >= 0 && col[k] < ncols, "csr_from_host: column index out of range");
      B2_REQUIRE(k == rowptr[r] || col[k] > col[k - 1], "csr_from_host: columns must be strictly ascending within a row");
    }
  }
  auto A = csr_alloc(c, nrows, ncols, nnz);
  B2_CUDA(cudaMemcpyAsync(A->rowptr.p, rowptr, sizeof(int) * ((size_t)nrows + 1), cudaMemcpyHostToDevice, c->stream));
  if (nnz) {
    B2_CUDA(cudaMemcpyAsync(A->col.p, col, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, c->stream));
    B2_CUDA(cudaMemcpyAsync(A->val.p, val, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, c->stream));
  }
  c->sync();
  A->plan();
  return A;
}

// ---------------------------------------------------------------- scale / add
namespace {
__global__ void __launch_bounds__(256) k_scale_cols(int64_t nnz, const int *__restrict__ col, const double *__restrict__ val, const double *__restrict__ d, double *out) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x) out[k] = val[k] * d[col[k]];
}
__global__ void __launch_bounds__(256) k_scale_cols_ghost(int64_t nnz, const int *__restrict__ col, const double *__restrict__ val, const double *__restrict__ d,
                                                          const double *ghost, int ncols, double *out) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x) {
    const int cl = col[k];
    out[k] = val[k] * (cl < ncols ? d[cl] : ghost[cl - ncols]);
  }
}
__global__ void __launch_bounds__(256) k_remap_ghost_cols(int64_t nnz, const int *__restrict__ col, int ncols, int dof, const int *__restrict__ map, int *out) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x) {
    const int cl = col[k];
    out[k] = cl < ncols ? cl : ncols + map[(cl - ncols) / dof] * dof + (cl - ncols) % dof;
  }
}
// rows of a row-partitioned DMDA matrix list their columns in GLOBAL order (ghost columns of lower ranks first), not in
// local-id order: the union merge below needs ascending ids, so the operands are sorted row by row first (short rows)
__global__ void __launch_bounds__(256) k_sort_rows(int nrows, const int *__restrict__ rowptr, int *col, double *val) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x) {
    const int b = rowptr[r], e = rowptr[r + 1];
    for (int i = b + 1; i < e; ++i) {
      const int cv = col[i];
      const double vv = val[i];
      int j = i;
      while (j > b && col[j - 1] > cv) { col[j] = col[j - 1]; val[j] = val[j - 1]; --j; }
      col[j] = cv; val[j] = vv;
    }
  }
}
// A + s*B on the union pattern: count, then fill (both matrices have ascending columns)
__global__ void __launch_bounds__(256) k_union_count(int nrows, const int *__restrict__ rpa, const int *__restrict__ ca, const int *__restrict__ rpb, const int *__restrict__ cb, int *cnt) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x) {
    int ka = rpa[r], ea = rpa[r + 1], kb = rpb[r], eb = rpb[r + 1], n = 0;
    while (ka < ea || kb < eb) {
      int a = ka < ea ? ca[ka] : 0x7fffffff, b = kb < eb ? cb[kb] : 0x7fffffff;
      if (a <= b) ka++;
      if (b <= a) kb++;
      n++;
    }
    cnt[r] = n;
  }
}
__global__ void __launch_bounds__(256) k_union_fill(int nrows, const int *__restrict__ rpa, const int *__restrict__ ca, const double *__restrict__ va,
                                                    const int *__restrict__ rpb, const int *__restrict__ cb, const double *__restrict__ vb, double s,
                                                    const int *__restrict__ rpc, int *cc, double *vc) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x) {
    int ka = rpa[r], ea = rpa[r + 1], kb = rpb[r], eb = rpb[r + 1], p = rpc[r];
    while (ka < ea || kb < eb) {
      int a = ka < ea ? ca[ka] : 0x7fffffff, b = kb < eb ? cb[kb] : 0x7fffffff;
      if (a == b) { cc[p] = a; vc[p++] = va[ka++] + s * vb[kb++]; }
      else if (a < b) { cc[p] = a; vc[p++] = va[ka++]; }
      else { cc[p] = b; vc[p++] = s * vb[kb++]; }
    }
  }
}
inline int grid_for(Ctx *c, int64_t n) {
  int64_t g = (n + 255) / 256;
  int64_t cap = (int64_t)c->num_sms * 8;
  return (int)std::max<int64_t>(1, std::min(g, cap));
}
} // namespace

std::shared_ptr<Csr> csr_scale_cols(const Csr &A, const double *d) {
  Ctx *c = A.ctx;
  if (A.halo) { // row-partitioned: the scale factors of the ghost columns travel through the matrix's own halo
    A.halo->begin(d, A.halo_dof);
    A.halo->end();
    const double *ghost = A.halo->ghost_now();
    auto C = csr_alloc(c, A.nrows, A.ncols, A.nnz);
    B2_CUDA(cudaMemcpyAsync(C->rowptr.p, A.rowptr.p, sizeof(int) * ((size_t)A.nrows + 1), cudaMemcpyDeviceToDevice, c->stream));
    B2_CUDA(cudaMemcpyAsync(C->col.p, A.col.p, sizeof(int) * (size_t)A.nnz, cudaMemcpyDeviceToDevice, c->stream));
    if (A.nnz) {
      LaunchScope ls(c, "setup");
      k_scale_cols_ghost<<<grid_for(c, A.nnz), 256, 0, c->stream>>>(A.nnz, A.col.p, A.val.p, d, ghost, A.ncols, C->val.p);
      check_launch("k_scale_cols_ghost");
    }
    c->sync();
    csr_copy_distribution(*C, A);
    C->plan();
    return C;
  }
  auto C = csr_alloc(c, A.nrows, A.ncols, A.nnz);
  B2_CUDA(cudaMemcpyAsync(C->rowptr.p, A.rowptr.p, sizeof(int) * ((size_t)A.nrows + 1), cudaMemcpyDeviceToDevice, c->stream));
  B2_CUDA(cudaMemcpyAsync(C->col.p, A.col.p, sizeof(int) * (size_t)A.nnz, cudaMemcpyDeviceToDevice, c->stream));
  if (A.nnz) {
    LaunchScope ls(c, "setup");
    k_scale_cols<<<grid_for(c, A.nnz), 256, 0, c->stream>>>(A.nnz, A.col.p, A.val.p, d, C->val.p);
    check_launch("k_scale_cols");
  }
  C->grid_M = A.grid_M; C->grid_N = A.grid_N; C->dof_r = A.dof_r; C->dof_c = A.dof_c;
  C->plan();
  return C;
}

std::shared_ptr<Csr> csr_add_scaled(const Csr &A, double s, const Csr &B) {
  Ctx *c = A.ctx;
  B2_REQUIRE(A.nrows == B.nrows && A.ncols == B.ncols, "csr_add_scaled: shape mismatch");
  B2_REQUIRE((A.halo == nullptr) == (B.halo == nullptr), "csr_add_scaled: one operand is row-partitioned and the other is not");
  // row-partitioned operands: the result lives in B's column space (its ghost set must contain A's: true for
  // A11 - A10 D^-1 A01, whose product stencil is a superset of A11's); A's ghost columns are renumbered into it
  DevBuf<int> a_cols, b_cols;
  DevBuf<double> a_vals, b_vals;
  const int *a_col = A.col.p, *b_col = B.col.p;
  const double *a_val = A.val.p, *b_val = B.val.p;
  if (A.halo) {
    B2_REQUIRE(A.halo_dof == B.halo_dof, "csr_add_scaled: column spaces with different dof per node");
    std::vector<int> map((size_t)A.halo->n_ghost + 1, 0);
    const auto &gb = B.halo->ghost_gnode;
    for (int t = 0; t < A.halo->n_ghost; ++t) {
      auto it = std::lower_bound(gb.begin(), gb.end(), A.halo->ghost_gnode[(size_t)t]);
      if (it == gb.end() || *it != A.halo->ghost_gnode[(size_t)t])
        throw Error(B200SP_ERR_UNSUPPORTED, "csr_add_scaled: the second operand's ghost columns do not contain the first operand's");
      map[(size_t)t] = (int)(it - gb.begin());
    }
    DevBuf<int> d_map(map.size());
    B2_CUDA(cudaMemcpyAsync(d_map.p, map.data(), sizeof(int) * map.size(), cudaMemcpyHostToDevice, c->stream));
    a_cols.alloc((size_t)A.nnz + 1); a_vals.alloc((size_t)A.nnz + 1);
    b_cols.alloc((size_t)B.nnz + 1); b_vals.alloc((size_t)B.nnz + 1);
    B2_CUDA(cudaMemcpyAsync(a_vals.p, A.val.p, sizeof(double) * (size_t)A.nnz, cudaMemcpyDeviceToDevice, c->stream));
    B2_CUDA(cudaMemcpyAsync(b_cols.p, B.col.p, sizeof(int) * (size_t)B.nnz, cudaMemcpyDeviceToDevice, c->stream));
    B2_CUDA(cudaMemcpyAsync(b_vals.p, B.val.p, sizeof(double) * (size_t)B.nnz, cudaMemcpyDeviceToDevice, c->stream));
    if (A.nnz) {
      LaunchScope ls(c, "setup");
      k_remap_ghost_cols<<<grid_for(c, A.nnz), 256, 0, c->stream>>>(A.nnz, A.col.p, A.ncols, A.halo_dof, d_map.p, a_cols.p);
      check_launch("k_remap_ghost_cols");
    }
    if (A.nrows) {
      LaunchScope ls(c, "setup");
      k_sort_rows<<<grid_for(c, A.nrows), 256, 0, c->stream>>>(A.nrows, A.rowptr.p, a_cols.p, a_vals.p);
      k_sort_rows<<<grid_for(c, A.nrows), 256, 0, c->stream>>>(A.nrows, B.rowptr.p, b_cols.p, b_vals.p);
      check_launch("k_sort_rows");
    }
    c->sync(); // d_map is freed on scope exit
    a_col = a_cols.p; a_val = a_vals.p; b_col = b_cols.p; b_val = b_vals.p;
  }
  DevBuf<int> cnt((size_t)A.nrows + 1);
  if (A.nrows) {
    LaunchScope ls(c, "setup");
    k_union_count<<<grid_for(c, A.nrows), 256, 0, c->stream>>>(A.nrows, A.rowptr.p, a_col, B.rowptr.p, b_col, cnt.p);
    check_launch("k_union_count");
  }
  DevBuf<int> rp((size_t)A.nrows + 1);
  int total = 0;
  exclusive_scan_i32(c, cnt.p, rp.p, A.nrows, &total);
  auto C = csr_alloc(c, A.nrows, A.ncols, total);
  B2_CUDA(cudaMemcpyAsync(C->rowptr.p, rp.p, sizeof(int) * ((size_t)A.nrows + 1), cudaMemcpyDeviceToDevice, c->stream));
  if (A.nrows) {
    LaunchScope ls(c, "setup");
    k_union_fill<<<grid_for(c, A.nrows), 256, 0, c->stream>>>(A.nrows, A.rowptr.p, a_col, a_val, B.rowptr.p, b_col, b_val, s, C->rowptr.p, C->col.p, C->val.p);
    check_launch("k_union_fill");
  }
  c->sync();
  C->grid_M = A.grid_M; C->grid_N = A.grid_N; C->dof_r = A.dof_r; C->dof_c = A.dof_c;
  if (B.halo) csr_copy_distribution(*C, B);
  C->plan();
  return C;
}

// ---------------------------------------------------------------- SpGEMM C = A*B (MatMatMult)
// Row-wise Gustavson, one thread per row of C, deterministic:
//   symbolic: the candidate columns of row i (concatenated B rows, each already ascending) are written to a
//             scratch segment sized by the upper bound sum_k len(B_k), insertion-sorted, made unique;
//   numeric : c_ij accumulates a_ik*b_kj in the order k appears in A's row (MatMatMultNumeric order),
//             product rounded then added, slot found by binary search in the sorted row.
namespace {
__global__ void __launch_bounds__(256) k_spgemm_ub(int nrows, const int *__restrict__ rpa, const int *__restrict__ ca, const int *__restrict__ rpb, int *ub) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x) {
    int n = 0;
    for (int k = rpa[r]; k < rpa[r + 1]; ++k) n += rpb[ca[k] + 1] - rpb[ca[k]];
    ub[r] = n;
  }
}
__global__ void __launch_bounds__(128) k_spgemm_symbolic(int nrows, const int *__restrict__ rpa, const int *__restrict__ ca, const int *__restrict__ rpb,
                                                         const int *__restrict__ cb, const int64_t *__restrict__ soff, int *scratch, int *cnt) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x) {
    int *s = scratch + soff[r];
    int n = 0;
    for (int k = rpa[r]; k < rpa[r + 1]; ++k)
      for (int t = rpb[ca[k]]; t < rpb[ca[k] + 1]; ++t) {
        // insertion into the sorted prefix (inputs arrive in ascending runs, so shifts are short)
        const int v = cb[t];
        int p = n;
        while (p > 0 && s[p - 1] > v) --p;
        if (p > 0 && s[p - 1] == v) continue;
        for (int q = n; q > p; --q) s[q] = s[q - 1];
        s[p] = v;
        ++n;
      }
    cnt[r] = n;
  }
}
__global__ void __launch_bounds__(128) k_spgemm_numeric(int nrows, const int *__restrict__ rpa, const int *__restrict__ ca, const double *__restrict__ va,
                                                        const int *__restrict__ rpb, const int *__restrict__ cb, const double *__restrict__ vb,
                                                        const int64_t *__restrict__ soff, const int *__restrict__ scratch,
                                                        const int *__restrict__ rpc, int *cc, double *vc) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x) {
    const int base = rpc[r], n = rpc[r + 1] - base;
    const int *s = scratch + soff[r];
    for (int t = 0; t < n; ++t) { cc[base + t] = s[t]; vc[base + t] = 0.0; }
    for (int k = rpa[r]; k < rpa[r + 1]; ++k) {
      const double a = va[k];
      for (int t = rpb[ca[k]]; t < rpb[ca[k] + 1]; ++t) {
        const int v = cb[t];
        int lo = 0, hi = n;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (cc[base + mid] < v) lo = mid + 1; else hi = mid; }
        vc[base + lo] += a * vb[t];
      }
    }
  }
}
} // namespace

std::shared_ptr<Csr> csr_matmat(const Csr &A, const Csr &B) {
  if (A.halo || B.halo) return csr_matmat_dist(A, B); // dist_spgemm.cu: fetch the ghost rows of B, multiply locally, build the result's halo
  B2_REQUIRE(A.ncols == B.nrows, "csr_matmat: inner dimensions differ");
  return spgemm_raw(A.ctx, A.nrows, A.rowptr.p, A.col.p, A.val.p, B.rowptr.p, B.col.p, B.val.p, B.ncols);
}

// distribution metadata of a matrix that shares `like`'s row partition and column space
void csr_copy_distribution(Csr &C, const Csr &like) {
  C.halo = like.halo; C.halo_dof = like.halo_dof; C.layout = like.layout;
  C.row_gstart = like.row_gstart; C.col_gstart = like.col_gstart;
  C.grid_M = like.grid_M; C.grid_N = like.grid_N; C.dof_r = like.dof_r; C.dof_c = like.dof_c;
}

std::shared_ptr<Csr> spgemm_raw(Ctx *c, int n, const int *rpa, const int *ca, const double *va, const int *rpb, const int *cb, const double *vb, int ncolsC) {
  DevBuf<int> ub((size_t)n + 1), ub_off((size_t)n + 1), cnt((size_t)n + 1), rp((size_t)n + 1);
  if (n) {
    LaunchScope ls(c, "setup");
    k_spgemm_ub<<<grid_for(c, n), 256, 0, c->stream>>>(n, rpa, ca, rpb, ub.p);
    check_launch("k_spgemm_ub");
  }
  // the scratch offsets can exceed 2^31 in principle: process in row chunks whose upper bound fits int32
  std::vector<int> h_ub((size_t)n);
  if (n) B2_CUDA(cudaMemcpyAsync(h_ub.data(), ub.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  c->sync();
  std::vector<int64_t> h_off((size_t)n + 1, 0);
  for (int r = 0; r < n; ++r) h_off[r + 1] = h_off[r] + h_ub[r];
  DevBuf<int64_t> soff((size_t)n + 1);
  B2_CUDA(cudaMemcpyAsync(soff.p, h_off.data(), sizeof(int64_t) * ((size_t)n + 1), cudaMemcpyHostToDevice, c->stream));
  DevBuf<int> scratch((size_t)h_off[n] + 1);
  if (n) {
    LaunchScope ls(c, "setup");
    k_spgemm_symbolic<<<std::max(1, std::min((n + 127) / 128, c->num_sms * 16)), 128, 0, c->stream>>>(n, rpa, ca, rpb, cb, soff.p, scratch.p, cnt.p);
    check_launch("k_spgemm_symbolic");
  }
  int total = 0;
  exclusive_scan_i32(c, cnt.p, rp.p, n, &total);
  auto C = csr_alloc(c, n, ncolsC, total);
  B2_CUDA(cudaMemcpyAsync(C->rowptr.p, rp.p, sizeof(int) * ((size_t)n + 1), cudaMemcpyDeviceToDevice, c->stream));
  if (n) {
    LaunchScope ls(c, "setup");
    k_spgemm_numeric<<<std::max(1, std::min((n + 127) / 128, c->num_sms * 16)), 128, 0, c->stream>>>(n, rpa, ca, va, rpb, cb, vb, soff.p, scratch.p,
                                                                                                      C->rowptr.p, C->col.p, C->val.p);
    check_launch("k_spgemm_numeric");
  }
  c->sync();
  C->plan();
  return C;
}

// ---------------------------------------------------------------- COO -> CSR: stable LSD radix sort + in-order sum
// key = (row << 32) | col, payload = position in the input (insertion order).  4-bit digits; every thread owns
// RS_ITEMS CONSECUTIVE keys of its tile, so "lower thread, then earlier item" is the input order and the
// scatter is stable.  After the sort equal keys are adjacent in insertion order; one thread per distinct key
// sums its duplicates in that order starting from +0.0 == MatSetValues(ADD_VALUES) in call order.
namespace {
constexpr int RS_ITEMS = 8, RS_THREADS = 256, RS_TILE = RS_ITEMS * RS_THREADS, RS_BINS = 16;

__global__ void __launch_bounds__(256) k_coo_keys(int64_t n, const int *__restrict__ row, const int *__restrict__ col, unsigned long long *keys, int *pay,
                                                  int nrows, int ncols, int *bad) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = row[i], c = col[i];
    if (r < 0 || r >= nrows || c < 0 || c >= ncols) *bad = 1;
    keys[i] = ((unsigned long long)(unsigned)r << 32) | (unsigned)c;
    pay[i] = (int)i;
  }
}
__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(int64_t n, const unsigned long long *__restrict__ keys, int shift, int *hist, int ntiles) {
  __shared__ int s_h[RS_BINS];
  if (threadIdx.x < RS_BINS) s_h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * RS_TILE + (int64_t)threadIdx.x * RS_ITEMS;
  int cnt[RS_BINS];
#pragma unroll
  for (int b = 0; b < RS_BINS; ++b) cnt[b] = 0;
#pragma unroll
  for (int k = 0; k < RS_ITEMS; ++k)
    if (base + k < n) {
      const int d = (int)((keys[base + k] >> shift) & (RS_BINS - 1));
#pragma unroll
      for (int b = 0; b < RS_BINS; ++b) cnt[b] += (d == b);
    }
#pragma unroll
  for (int b = 0; b < RS_BINS; ++b) {
    int v = cnt[b];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_h[b], v);
  }
  __syncthreads();
  if (threadIdx.x < RS_BINS) hist[(size_t)threadIdx.x * ntiles + blockIdx.x] = s_h[threadIdx.x]; // bin-major for the global scan
}
__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(int64_t n, const unsigned long long *__restrict__ keys, const int *__restrict__ pay, int shift,
                                                           const int *__restrict__ offs, int ntiles, unsigned long long *keys_out, int *pay_out) {
  __shared__ int s_cnt[RS_BINS][RS_THREADS + 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t base = (int64_t)blockIdx.x * RS_TILE + (int64_t)tid * RS_ITEMS;
  unsigned long long kk[RS_ITEMS];
  int pp[RS_ITEMS], dd[RS_ITEMS];
  int cnt[RS_BINS];
#pragma unroll
  for (int b = 0; b < RS_BINS; ++b) cnt[b] = 0;
#pragma unroll
  for (int k = 0; k < RS_ITEMS; ++k) {
    dd[k] = -1;
    if (base + k < n) {
      kk[k] = keys[base + k];
      pp[k] = pay[base + k];
      dd[k] = (int)((kk[k] >> shift) & (RS_BINS - 1));
#pragma unroll
      for (int b = 0; b < RS_BINS; ++b) cnt[b] += (dd[k] == b);
    }
  }
#pragma unroll
  for (int b = 0; b < RS_BINS; ++b) s_cnt[b][tid] = cnt[b];
  __syncthreads();
  // exclusive scan over the 256 threads for each bin: warp w scans bins 2w and 2w+1
  for (int b = warp * 2; b < warp * 2 + 2; ++b) {
    int carry = 0;
    for (int seg = 0; seg < RS_THREADS; seg += 32) {
      const int v = s_cnt[b][seg + lane];
      int incl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += u;
      }
      s_cnt[b][seg + lane] = carry + incl - v;
      carry += __shfl_sync(FULL, incl, 31);
    }
  }
  __syncthreads();
  int run[RS_BINS];
#pragma unroll
  for (int b = 0; b < RS_BINS; ++b) run[b] = offs[(size_t)b * ntiles + blockIdx.x] + s_cnt[b][tid];
#pragma unroll
  for (int k = 0; k < RS_ITEMS; ++k)
    if (dd[k] >= 0) {
      int dest = 0;
#pragma unroll
      for (int b = 0; b < RS_BINS; ++b)
        if (dd[k] == b) dest = run[b]++;
      keys_out[dest] = kk[k];
      pay_out[dest] = pp[k];
    }
}
__global__ void __launch_bounds__(256) k_coo_heads(int64_t n, const unsigned long long *__restrict__ keys, int *head) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}
// one thread per sorted entry that is a segment head: sum the run of equal keys in order, write col/val, and mark row starts
__global__ void __launch_bounds__(256) k_coo_reduce(int64_t n, const unsigned long long *__restrict__ keys, const int *__restrict__ pay,
                                                    const int *__restrict__ head, const int *__restrict__ upos, const double *__restrict__ val,
                                                    int *col_out, double *val_out, int *rowcnt) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (!head[i]) continue;
    const unsigned long long k = keys[i];
    double s = 0.0;
    for (int64_t j = i; j < n && keys[j] == k; ++j) s += val[pay[j]];
    const int u = upos[i];
    col_out[u] = (int)(k & 0xffffffffu);
    val_out[u] = s;
    atomicAdd(&rowcnt[(int)(k >> 32)], 1);
  }
}
int bits_for(int v) { int b = 0; while ((1LL << b) < (long long)v) ++b; return b < 1 ? 1 : b; }
} // namespace

std::shared_ptr<Csr> csr_from_coo_device(Ctx *c, int nrows, int ncols, int64_t n, const int *d_row, const int *d_col, const double *d_val) {
  B2_REQUIRE(nrows >= 0 && ncols >= 0 && n >= 0 && n < (1LL << 31), "csr_from_coo: bad sizes");
  DevBuf<unsigned long long> k0((size_t)n + 1), k1((size_t)n + 1);
  DevBuf<int> p0((size_t)n + 1), p1((size_t)n + 1), bad(1);
  bad.zero(c->stream);
  if (n) {
    LaunchScope ls(c, "setup");
    k_coo_keys<<<grid_for(c, n), 256, 0, c->stream>>>(n, d_row, d_col, k0.p, p0.p, nrows, ncols, bad.p);
    check_launch("k_coo_keys");
  }
  int h_bad = 0;
  B2_CUDA(cudaMemcpyAsync(&h_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  c->sync();
  B2_REQUIRE(!h_bad, "csr_from_coo: row or column index out of range");
  const int ntiles = (int)((n + RS_TILE - 1) / RS_TILE);
  DevBuf<int> hist((size_t)RS_BINS * (ntiles > 0 ? ntiles : 1) + 1), offs((size_t)RS_BINS * (ntiles > 0 ? ntiles : 1) + 1);
  unsigned long long *ka = k0.p, *kb = k1.p;
  int *pa = p0.p, *pb = p1.p;
  // LSD: the column bits first (low word), then the row bits (high word)
  std::vector<int> shifts;
  for (int s = 0; s < bits_for(ncols); s += 4) shifts.push_back(s);
  for (int s = 0; s < bits_for(nrows); s += 4) shifts.push_back(32 + s);
  for (int shift : shifts) {
    if (!n) break;
    {
      LaunchScope ls(c, "setup");
      k_rs_hist<<<ntiles, RS_THREADS, 0, c->stream>>>(n, ka, shift, hist.p, ntiles);
      check_launch("k_rs_hist");
    }
    exclusive_scan_i32(c, hist.p, offs.p, (int64_t)RS_BINS * ntiles, nullptr);
    {
      LaunchScope ls(c, "setup");
      k_rs_scatter<<<ntiles, RS_THREADS, 0, c->stream>>>(n, ka, pa, shift, offs.p, ntiles, kb, pb);
      check_launch("k_rs_scatter");
    }
    std::swap(ka, kb);
    std::swap(pa, pb);
  }
  DevBuf<int> head((size_t)n + 1), upos((size_t)n + 2);
  int nuniq = 0;
  if (n) {
    LaunchScope ls(c, "setup");
    k_coo_heads<<<grid_for(c, n), 256, 0, c->stream>>>(n, ka, head.p);
    check_launch("k_coo_heads");
  }
  exclusive_scan_i32(c, head.p, upos.p, n, &nuniq);
  auto A = csr_alloc(c, nrows, ncols, nuniq);
  DevBuf<int> rowcnt((size_t)nrows + 1);
  rowcnt.zero(c->stream);
  if (n) {
    LaunchScope ls(c, "setup");
    k_coo_reduce<<<grid_for(c, n), 256, 0, c->stream>>>(n, ka, pa, head.p, upos.p, d_val, A->col.p, A->val.p, rowcnt.p);
    check_launch("k_coo_reduce");
  }
  exclusive_scan_i32(c, rowcnt.p, A->rowptr.p, nrows, nullptr);
  c->sync();
  A->plan();
  return A;
}

std::shared_ptr<Csr> csr_from_coo_host(Ctx *c, int nrows, int ncols, int64_t n, const int *row, const int *col, const double *val) {
  DevBuf<int> d_row((size_t)n + 1), d_col((size_t)n + 1);
  DevBuf<double> d_val((size_t)n + 1);
  if (n) {
    B2_CUDA(cudaMemcpyAsync(d_row.p, row, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    B2_CUDA(cudaMemcpyAsync(d_col.p, col, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    B2_CUDA(cudaMemcpyAsync(d_val.p, val, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  }
  return csr_from_coo_device(c, nrows, ncols, n, d_row.p, d_col.p, d_val.p);
}

// explicit transpose = COO (col, row, val) through the same stable sort: rows of A^T come out with ascending
// columns (= ascending original rows), the order MatTranspose produces
namespace {
__global__ void __launch_bounds__(256) k_expand_rows(int nrows, const int *__restrict__ rowptr, int *rows) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x)
    for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) rows[k] = r;
}
} // namespace
std::shared_ptr<Csr> csr_transpose(const Csr &A) {
  if (A.halo) throw Error(B200SP_ERR_UNSUPPORTED, "csr_transpose: row-partitioned matrices are not supported by this routine (ghost columns)");

  Ctx *c = A.ctx;
  DevBuf<int> rows((size_t)A.nnz + 1);
  if (A.nrows) {
    LaunchScope ls(c, "setup");
    k_expand_rows<<<grid_for(c, A.nrows), 256, 0, c->stream>>>(A.nrows, A.rowptr.p, rows.p);
    check_launch("k_expand_rows");
  }
  return csr_from_coo_device(c, A.ncols, A.nrows, A.nnz, A.col.p, rows.p, A.val.p);
}

// ---------------------------------------------------------------- MatCreateSubMatrix for strided fields (PCFIELDSPLIT
// on a monolithic matrix with block size bs: the reference's case, KSPSetOperators(A,A) with the DMDA's bs = 2)
namespace {
struct FieldMap { int bs; int split[8]; int pos[8]; int nf[2]; };
__global__ void __launch_bounds__(256) k_extract_count(int nrows_src, FieldMap fm, int rs, int cs, const int *__restrict__ rowptr, const int *__restrict__ col, int *cnt) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows_src; r += gridDim.x * blockDim.x) {
    const int f = r % fm.bs;
    if (fm.split[f] != rs) continue;
    int n = 0;
    for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) n += fm.split[col[k] % fm.bs] == cs;
    cnt[(r / fm.bs) * fm.nf[rs] + fm.pos[f]] = n;
  }
}
__global__ void __launch_bounds__(256) k_extract_fill(int nrows_src, FieldMap fm, int rs, int cs, const int *__restrict__ rowptr, const int *__restrict__ col,
                                                      const double *__restrict__ val, const int *__restrict__ rp, int *cj, double *va) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows_src; r += gridDim.x * blockDim.x) {
    const int f = r % fm.bs;
    if (fm.split[f] != rs) continue;
    int p = rp[(r / fm.bs) * fm.nf[rs] + fm.pos[f]];
    for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) {
      const int c = col[k], fc = c % fm.bs;
      if (fm.split[fc] != cs) continue;
      cj[p] = (c / fm.bs) * fm.nf[cs] + fm.pos[fc];
      va[p++] = val[k];
    }
  }
}
} // namespace

// fields_of_split1: the fields (ascending) that form split 1; every other field goes to split 0
std::shared_ptr<Csr> csr_extract_fields(const Csr &A, int bs, const std::vector<int> &split_of_field, int rs, int cs) {
  Ctx *c = A.ctx;
  B2_REQUIRE(bs >= 2 && bs <= 8 && A.nrows % bs == 0 && A.ncols % bs == 0 && (int)split_of_field.size() == bs, "extract_fields: bad block size");
  B2_REQUIRE(!A.halo, "extract_fields: row-partitioned matrices are not supported");
  FieldMap fm;
  fm.bs = bs; fm.nf[0] = fm.nf[1] = 0;
  for (int f = 0; f < bs; ++f) { fm.split[f] = split_of_field[(size_t)f]; fm.pos[f] = fm.nf[fm.split[f]]++; }
  B2_REQUIRE(fm.nf[0] > 0 && fm.nf[1] > 0, "extract_fields: both splits need at least one field");
  const int nr = A.nrows / bs * fm.nf[rs], nc = A.ncols / bs * fm.nf[cs];
  DevBuf<int> cnt((size_t)nr + 1), rp((size_t)nr + 1);
  {
    LaunchScope ls(c, "setup");
    k_extract_count<<<grid_for(c, A.nrows), 256, 0, c->stream>>>(A.nrows, fm, rs, cs, A.rowptr.p, A.col.p, cnt.p);
    check_launch("k_extract_count");
  }
  int nnz = 0;
  exclusive_scan_i32(c, cnt.p, rp.p, nr, &nnz);
  auto S = csr_alloc(c, nr, nc, nnz);
  B2_CUDA(cudaMemcpyAsync(S->rowptr.p, rp.p, sizeof(int) * ((size_t)nr + 1), cudaMemcpyDeviceToDevice, c->stream));
  {
    LaunchScope ls(c, "setup");
    k_extract_fill<<<grid_for(c, A.nrows), 256, 0, c->stream>>>(A.nrows, fm, rs, cs, A.rowptr.p, A.col.p, A.val.p, S->rowptr.p, S->col.p, S->val.p);
    check_launch("k_extract_fill");
  }
  c->sync();
  S->plan();
  return S;
}

// ---------------------------------------------------------------- small dense inverse (coarsest MG level)
namespace {
__global__ void __launch_bounds__(256) k_csr_to_dense_aug(int n, const int *__restrict__ rowptr, const int *__restrict__ col, const double *__restrict__ val, double *W /* n x 2n */) {
  for (int r = blockIdx.x; r < n; r += gridDim.x) {
    for (int k = rowptr[r] + threadIdx.x; k < rowptr[r + 1]; k += blockDim.x) W[(size_t)r * 2 * n + col[k]] = val[k];
    if (threadIdx.x == 0) W[(size_t)r * 2 * n + n + r] = 1.0;
  }
}
// Gauss-Jordan with partial pivoting on the augmented matrix [A | I], single CTA (n <= a few thousand).
__global__ void __launch_bounds__(1024) k_gauss_jordan(int n, double *W, int *info) {
  __shared__ int s_piv;
  __shared__ double s_best[32];
  __shared__ int s_bidx[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int w2 = 2 * n;
  for (int k = 0; k < n; ++k) {
    // pivot search in column k, rows k..n-1
    double best = -1.0;
    int bidx = k;
    for (int r = k + tid; r < n; r += blockDim.x) {
      double a = fabs(W[(size_t)r * w2 + k]);
      if (a > best) { best = a; bidx = r; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      double ob = __shfl_xor_sync(FULL, best, o);
      int oi = __shfl_xor_sync(FULL, bidx, o);
      if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
    }
    if (lane == 0) { s_best[warp] = best; s_bidx[warp] = bidx; }
    __syncthreads();
    if (tid == 0) {
      double b = s_best[0];
      int bi = s_bidx[0];
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
        if (s_best[w] > b || (s_best[w] == b && s_bidx[w] < bi)) { b = s_best[w]; bi = s_bidx[w]; }
      s_piv = bi;
      if (!(b > 0.0)) *info = k + 1;
    }
    __syncthreads();
    const int p = s_piv;
    if (p != k)
      for (int j = tid; j < w2; j += blockDim.x) {
        double t = W[(size_t)k * w2 + j];
        W[(size_t)k * w2 + j] = W[(size_t)p * w2 + j];
        W[(size_t)p * w2 + j] = t;
      }
    __syncthreads();
    const double inv = 1.0 / W[(size_t)k * w2 + k];
    __syncthreads();
    for (int j = tid; j < w2; j += blockDim.x) W[(size_t)k * w2 + j] *= inv;
    __syncthreads();
    // eliminate column k from every other row: warp per row
    for (int r = warp; r < n; r += (blockDim.x >> 5)) {
      if (r == k) continue;
      const double f = W[(size_t)r * w2 + k];
      if (f != 0.0)
        for (int j = lane; j < w2; j += 32)
          if (j != k) W[(size_t)r * w2 + j] -= f * W[(size_t)k * w2 + j];
    }
    __syncthreads();
    for (int r = tid; r < n; r += blockDim.x)
      if (r != k) W[(size_t)r * w2 + k] = 0.0;
    __syncthreads();
  }
}
__global__ void __launch_bounds__(256) k_extract_inverse(int n, const double *__restrict__ W, double *__restrict__ Ainv) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < (int64_t)n * n; t += (int64_t)gridDim.x * blockDim.x) {
    int r = (int)(t / n), j = (int)(t % n);
    Ainv[t] = W[(size_t)r * 2 * n + n + j];
  }
}
// y = Ainv x, warp per row
__global__ void __launch_bounds__(256) k_dense_matvec(int n, const double *__restrict__ Ainv, const double *__restrict__ x, double *y) {
  const int lane = threadIdx.x & 31;
  for (int r = blockIdx.x * 8 + (threadIdx.x >> 5); r < n; r += gridDim.x * 8) {
    double s = 0.0;
    for (int j = lane; j < n; j += 32) s = fma(Ainv[(size_t)r * n + j], x[j], s);
    s = warp_sum(s);
    if (lane == 0) y[r] = s;
  }
}
} // namespace

void dense_inverse_from_csr(const Csr &A, double *Ainv) {
  if (A.halo) throw Error(B200SP_ERR_UNSUPPORTED, "dense_inverse_from_csr: row-partitioned matrices are not supported by this routine (ghost columns)");

  Ctx *c = A.ctx;
  const int n = A.nrows;
  B2_REQUIRE(n == A.ncols && n > 0 && n <= 8192, "dense_inverse_from_csr: matrix must be square with n <= 8192");
  DevBuf<double> W((size_t)n * 2 * n);
  DevBuf<int> info(1);
  W.zero(c->stream);
  info.zero(c->stream);
  {
    LaunchScope ls(c, "setup");
    k_csr_to_dense_aug<<<std::min(n, c->num_sms * 8), 64, 0, c->stream>>>(n, A.rowptr.p, A.col.p, A.val.p, W.p);
    check_launch("k_csr_to_dense_aug");
  }
  {
    LaunchScope ls(c, "setup");
    k_gauss_jordan<<<1, 1024, 0, c->stream>>>(n, W.p, info.p);
    check_launch("k_gauss_jordan");
  }
  {
    LaunchScope ls(c, "setup");
    k_extract_inverse<<<grid_for(c, (int64_t)n * n), 256, 0, c->stream>>>(n, W.p, Ainv);
    check_launch("k_extract_inverse");
  }
  int h_info = 0;
  B2_CUDA(cudaMemcpyAsync(&h_info, info.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  c->sync();
  if (h_info) throw Error(B200SP_ERR_ARG, "dense_inverse_from_csr: singular matrix at pivot " + std::to_string(h_info - 1));
}

void dense_matvec(Ctx *c, int n, const double *Ainv, const double *x, double *y) {
  LaunchScope ls(c, "coarse");
  int grid = std::min((n + 7) / 8, c->num_sms * 4);
  k_dense_matvec<<<grid, 256, 0, c->stream>>>(n, Ainv, x, y);
  check_launch("k_dense_matvec");
}

} // namespace b200sp
