// kernels_spmv.cu -- CSR SpMV for A, B, B^T, C, P, R (replaces MatMult_SeqAIJ / MatMultAdd reached from
// KSPSolve and PCApply; reference call site src/SaddlePointProblem.c:70).
//
// Algorithmic bytes per launch: 12*nnz + 4*(rows+1) + 8*rows + 8*cols.
// Kernel choice from the row-length histogram (Csr::plan):
//   SPMV_STREAM  short rows (every 32-row group fits the per-warp shared tile): a warp owns 32 consecutive
//                rows = ONE contiguous segment of val/col.  Phase 1 streams that segment with coalesced
//                128-bit loads (L1 no-allocate), gathers x through L1, and parks the products in the warp's
//                shared tile.  Phase 2: lane l sums row l's products in CSR order -> the result is
//                bit-identical to the sequential MatMult_SeqAIJ loop (product rounded, then added).
//                No block barrier, only __syncwarp; loads are independent of row boundaries so lanes never
//                idle on short rows (8/12/18 nnz for A, 4/6/9 for B^T).
//   SPMV_VECTOR  medium rows: LPR lanes per row (2..32), shuffle-tree reduction.
//   SPMV_BLOCK   very long rows (dense constraint rows): one CTA per row.
// Epilogue (fused): y = beta_z*z + alpha*(A x)  covers MatMult, MatMultAdd and the residual b - A x.
#include "dev.cuh"
#include "dist.h"
#include <type_traits>

namespace b200sp {

namespace {

constexpr int STREAM_WARPS = 8;           // 256 threads per CTA
constexpr int STREAM_MAX_GROUP_NNZ = 1152; // per-warp tile limit (9 KB): up to 36 nnz/row on average

__device__ __forceinline__ double epilogue(double s, double alpha, const double *z, double beta_z, int r) {
  double v = alpha * s;
  if (z) v = beta_z * z[r] + v;
  return v;
}

__global__ void __launch_bounds__(256) k_spmv_stream(int nrows, const int *__restrict__ rowptr, const int *__restrict__ col,
                                                     const double *__restrict__ val, const double *__restrict__ x, double *y,
                                                     double alpha, const double *z, double beta_z, int tile_elems) {
  extern __shared__ double s_prod[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double *prod = s_prod + (size_t)warp * tile_elems;
  const int ngroups = (nrows + 31) >> 5;
  const int gstride = gridDim.x * STREAM_WARPS;
  for (int g = blockIdx.x * STREAM_WARPS + warp; g < ngroups; g += gstride) {
    const int r = (g << 5) + lane;
    const int rs = rowptr[r < nrows ? r : nrows];
    const int re = rowptr[r + 1 < nrows ? r + 1 : nrows];
    const int s = __shfl_sync(FULL, rs, 0);
    const int e = __shfl_sync(FULL, re, 31);
    const int s_al = s & ~1; // 16-byte alignment of the val stream (col stream is then 8-byte aligned)
    // phase 1: stream the segment [s_al, e), two entries per lane per step
#pragma unroll 4
    for (int base = s_al + 2 * lane; base < e; base += 64) {
      const double2 v = ld_stream_f64x2(val + base);
      const int2 c = ld_stream_s32x2(col + base);
      double2 p;
      p.x = v.x * __ldg(x + c.x);
      p.y = v.y * __ldg(x + c.y);
      *reinterpret_cast<double2 *>(prod + (base - s_al)) = p;
    }
    __syncwarp();
    // phase 2: one row per lane, CSR order
    if (r < nrows) {
      double sum = 0.0;
      for (int k = rs - s_al; k < re - s_al; ++k) sum += prod[k];
      y[r] = epilogue(sum, alpha, z, beta_z, r);
    }
    __syncwarp();
  }
}

template <int LPR>
__global__ void __launch_bounds__(256) k_spmv_vector(int nrows, const int *__restrict__ rowptr, const int *__restrict__ col,
                                                     const double *__restrict__ val, const double *__restrict__ x, double *y,
                                                     double alpha, const double *z, double beta_z) {
  constexpr int rows_per_cta = 256 / LPR;
  constexpr int rows_per_warp = 32 / LPR;
  const int sub = threadIdx.x % LPR;
  const int lane = threadIdx.x & 31;
  // the loop condition is warp-uniform (first row of the warp) so the shuffles below see every lane
  for (int rw = blockIdx.x * rows_per_cta + (threadIdx.x >> 5) * rows_per_warp; rw < nrows; rw += gridDim.x * rows_per_cta) {
    const int r = rw + lane / LPR;
    double sum = 0.0;
    if (r < nrows) {
      const int rs = rowptr[r], re = rowptr[r + 1];
      for (int k = rs + sub; k < re; k += LPR) sum += ld_stream_f64(val + k) * __ldg(x + ld_stream_s32(col + k));
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(FULL, sum, o);
    if (r < nrows && sub == 0) y[r] = epilogue(sum, alpha, z, beta_z, r);
  }
}

__global__ void __launch_bounds__(256) k_spmv_block(int nrows, const int *__restrict__ rowptr, const int *__restrict__ col,
                                                    const double *__restrict__ val, const double *__restrict__ x, double *y,
                                                    double alpha, const double *z, double beta_z) {
  __shared__ double s_w[8];
  for (int r = blockIdx.x; r < nrows; r += gridDim.x) {
    const int rs = rowptr[r], re = rowptr[r + 1];
    double sum = 0.0;
    for (int k = rs + threadIdx.x; k < re; k += 256) sum += ld_stream_f64(val + k) * __ldg(x + ld_stream_s32(col + k));
    sum = warp_sum(sum);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += s_w[w];
      y[r] = epilogue(t, alpha, z, beta_z, r);
    }
    __syncthreads();
  }
}

// per 32-row group nnz (max) + row-length histogram, one pass over rowptr
__global__ void __launch_bounds__(256) k_row_stats(int nrows, const int *__restrict__ rowptr, int *max_row, int *max_group, unsigned long long *hist) {
  __shared__ unsigned int s_hist[14];
  if (threadIdx.x < 14) s_hist[threadIdx.x] = 0;
  __syncthreads();
  int mr = 0, mg = 0;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x) {
    const int len = rowptr[r + 1] - rowptr[r];
    mr = max(mr, len);
    int bin = len <= 2 ? len : (len > 2048 ? 13 : 1 + (32 - __clz(len - 1))); // 3-4 -> 3, 5-8 -> 4, ..., 1025-2048 -> 12, >2048 -> 13
    atomicAdd(&s_hist[bin], 1u);
    if ((r & 31) == 0) {
      const int rend = r + 32 < nrows ? r + 32 : nrows;
      mg = max(mg, rowptr[rend] - (rowptr[r] & ~1));
    }
  }
  mr = warp_max(mr);
  mg = warp_max(mg);
  if ((threadIdx.x & 31) == 0) { atomicMax(max_row, mr); atomicMax(max_group, mg); }
  __syncthreads();
  if (threadIdx.x < 14 && s_hist[threadIdx.x]) atomicAdd(hist + threadIdx.x, (unsigned long long)s_hist[threadIdx.x]);
}

__global__ void __launch_bounds__(256) k_get_diag(int nrows, const int *__restrict__ rowptr, const int *__restrict__ col, const double *__restrict__ val, double *d) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x) {
    double v = 0.0;
    int lo = rowptr[r], hi = rowptr[r + 1];
    while (lo < hi) { // columns ascending: binary search
      int mid = (lo + hi) >> 1;
      int c = col[mid];
      if (c < r) lo = mid + 1; else hi = mid;
    }
    if (lo < rowptr[r + 1] && col[lo] == r) v = val[lo];
    d[r] = v;
  }
}

// MatZeroRowsColumns / MatZeroRows / zero columns: one thread per row, flags in a byte map
__global__ void __launch_bounds__(256) k_zero_rows_cols(int nrows, const int *__restrict__ rowptr, const int *__restrict__ col, double *val,
                                                        const unsigned char *__restrict__ rowflag, const unsigned char *__restrict__ colflag,
                                                        double diag, int set_diag) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x) {
    const bool rz = rowflag && rowflag[r];
    for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) {
      const int c = col[k];
      if (rz) val[k] = (set_diag && c == r) ? diag : 0.0;
      else if (colflag && colflag[c]) val[k] = 0.0;
    }
  }
}
// y[off_rows[k]] += alpha * (row k of the off-diagonal block) . ghost values   (the second half of MatMult_MPIAIJ)
__global__ void __launch_bounds__(128) k_spmv_offdiag(int n, const int *__restrict__ rowptr, const int *__restrict__ col, const double *__restrict__ val,
                                                      const int *__restrict__ off_rows, const double *__restrict__ ghost, double *y, double alpha) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int t = rowptr[k]; t < rowptr[k + 1]; ++t) s += val[t] * ghost[col[t]];
    y[off_rows[k]] += alpha * s;
  }
}
// zero the off-diagonal block: whole rows whose local row is flagged, and entries whose ghost column is flagged
__global__ void __launch_bounds__(128) k_zero_offdiag(int n, const int *__restrict__ rowptr, const int *__restrict__ col, double *val,
                                                      const int *__restrict__ off_rows, const unsigned char *__restrict__ rowflag,
                                                      const double *__restrict__ ghostflag) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const bool rz = rowflag && rowflag[off_rows[k]];
    for (int t = rowptr[k]; t < rowptr[k + 1]; ++t)
      if (rz || (ghostflag && ghostflag[col[t]] != 0.0)) val[t] = 0.0;
  }
}
__global__ void k_flag_to_double(int n, const unsigned char *flag, double *out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = flag[i] ? 1.0 : 0.0;
}
__global__ void k_mark(int n, const int *idx, unsigned char *flag) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) flag[idx[i]] = 1;
}

} // namespace

void Csr::plan() {
  DevBuf<int> d_max(2);
  DevBuf<unsigned long long> d_hist(14);
  d_max.zero(ctx->stream);
  d_hist.zero(ctx->stream);
  if (nrows > 0) {
    LaunchScope ls(ctx, "setup");
    int grid = (nrows + 255) / 256;
    if (grid > ctx->num_sms * 8) grid = ctx->num_sms * 8;
    k_row_stats<<<grid, 256, 0, ctx->stream>>>(nrows, rowptr.p, d_max.p, d_max.p + 1, d_hist.p);
    check_launch("k_row_stats");
  }
  int h_max[2];
  unsigned long long h_hist[14];
  B2_CUDA(cudaMemcpyAsync(h_max, d_max.p, sizeof(h_max), cudaMemcpyDeviceToHost, ctx->stream));
  B2_CUDA(cudaMemcpyAsync(h_hist, d_hist.p, sizeof(h_hist), cudaMemcpyDeviceToHost, ctx->stream));
  ctx->sync();
  max_row_nnz = h_max[0];
  max_group_nnz = h_max[1] + 2;
  for (int i = 0; i < 14; ++i) hist[i] = (int64_t)h_hist[i];
  const double mean = nrows ? (double)nnz / nrows : 0.0;
  if (max_group_nnz <= STREAM_MAX_GROUP_NNZ) kernel = SPMV_TMA; // short rows: TMA-staged thread-per-row (SPMV_STREAM is the non-TMA alternative)
  else if (mean >= 2048.0) kernel = SPMV_BLOCK;
  else kernel = SPMV_VECTOR;
  lanes_per_row = mean <= 2 ? 2 : mean <= 4 ? 4 : mean <= 8 ? 8 : mean <= 16 ? 16 : 32;
}

static void csr_spmv_local(const Csr &A, const double *x, double *y, double alpha, const double *z, double beta_z);

// Distributed MatMult (MatMult_MPIAIJ): start the halo exchange of x, multiply the diagonal block while the
// ghosts travel (separate stream), then add the off-diagonal block's contribution on the boundary rows.
void csr_spmv(const Csr &A, const double *x, double *y, double alpha, const double *z, double beta_z) {
  Ctx *c = A.ctx;
  const bool dist = A.halo && c->dcomm;
  if (dist) A.halo->begin(x, A.halo_dof);
  csr_spmv_local(A, x, y, alpha, z, beta_z);
  if (dist) {
    A.halo->end();
    if (A.off && A.off->nrows > 0) {
      LaunchScope ls(c, "spmv:offdiag");
      const int n = A.off->nrows;
      k_spmv_offdiag<<<(n + 127) / 128, 128, 0, c->stream>>>(n, A.off->rowptr.p, A.off->col.p, A.off->val.p, A.off_rows.p, A.halo->ghost.p, y, alpha);
      check_launch("k_spmv_offdiag");
    }
  }
}

static void csr_spmv_local(const Csr &A, const double *x, double *y, double alpha, const double *z, double beta_z) {
  Ctx *c = A.ctx;
  if (A.nrows <= 0) return;
  LaunchScope ls(c, A.tag.c_str());
  if (A.kernel == SPMV_TMA && csr_spmv_tma(A, x, y, alpha, z, beta_z)) return;
  if (A.kernel == SPMV_STREAM || A.kernel == SPMV_TMA) {
    int tile = (A.max_group_nnz + 1) & ~1;
    size_t smem = (size_t)tile * sizeof(double) * STREAM_WARPS;
    static bool attr_set = false;
    if (!attr_set) {
      B2_CUDA(cudaFuncSetAttribute(k_spmv_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, STREAM_MAX_GROUP_NNZ * 8 * STREAM_WARPS + 64));
      attr_set = true;
    }
    int per_sm = (int)((220 * 1024) / (smem + 1024));
    if (per_sm > 8) per_sm = 8;
    if (per_sm < 1) per_sm = 1;
    int ngroups = (A.nrows + 31) / 32;
    int grid = (ngroups + STREAM_WARPS - 1) / STREAM_WARPS;
    if (grid > c->num_sms * per_sm) grid = c->num_sms * per_sm;
    k_spmv_stream<<<grid, 256, smem, c->stream>>>(A.nrows, A.rowptr.p, A.col.p, A.val.p, x, y, alpha, z, beta_z, tile);
    check_launch("k_spmv_stream");
  } else if (A.kernel == SPMV_BLOCK) {
    int grid = A.nrows < c->num_sms * 8 ? A.nrows : c->num_sms * 8;
    k_spmv_block<<<grid, 256, 0, c->stream>>>(A.nrows, A.rowptr.p, A.col.p, A.val.p, x, y, alpha, z, beta_z);
    check_launch("k_spmv_block");
  } else {
    auto launch = [&](auto lpr_tag) {
      constexpr int LPR = decltype(lpr_tag)::value;
      int rows_per_cta = 256 / LPR;
      int grid = (A.nrows + rows_per_cta - 1) / rows_per_cta;
      if (grid > c->num_sms * 8) grid = c->num_sms * 8;
      k_spmv_vector<LPR><<<grid, 256, 0, c->stream>>>(A.nrows, A.rowptr.p, A.col.p, A.val.p, x, y, alpha, z, beta_z);
    };
    switch (A.lanes_per_row) {
    case 2: launch(std::integral_constant<int, 2>()); break;
    case 4: launch(std::integral_constant<int, 4>()); break;
    case 8: launch(std::integral_constant<int, 8>()); break;
    case 16: launch(std::integral_constant<int, 16>()); break;
    default: launch(std::integral_constant<int, 32>()); break;
    }
    check_launch("k_spmv_vector");
  }
}

void csr_get_diagonal(const Csr &A, double *d) {
  if (A.nrows <= 0) return;
  LaunchScope ls(A.ctx, "setup");
  int grid = (A.nrows + 255) / 256;
  k_get_diag<<<grid, 256, 0, A.ctx->stream>>>(A.nrows, A.rowptr.p, A.col.p, A.val.p, d);
  check_launch("k_get_diag");
}

void csr_zero_rows_cols(Csr &A, int n, const int *rows_host, double diag, bool do_rows, bool do_cols, bool set_diag) {
  Ctx *c = A.ctx;
  const bool dist = A.halo && c->dcomm;
  if ((n <= 0 && !dist) || A.nrows <= 0) return; // collective when distributed: ranks without ids still exchange flags
  DevBuf<int> d_idx((size_t)n + 1);
  if (n > 0) B2_CUDA(cudaMemcpyAsync(d_idx.p, rows_host, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  DevBuf<unsigned char> rowflag, colflag;
  int g = (n + 255) / 256;
  if (g < 1) g = 1;
  if (do_rows) {
    rowflag.alloc((size_t)A.nrows);
    rowflag.zero(c->stream);
    LaunchScope ls(c, "setup");
    k_mark<<<g, 256, 0, c->stream>>>(n, d_idx.p, rowflag.p);
  }
  if (do_cols) {
    colflag.alloc((size_t)A.ncols);
    colflag.zero(c->stream);
    LaunchScope ls(c, "setup");
    k_mark<<<g, 256, 0, c->stream>>>(n, d_idx.p, colflag.p);
  }
  if (dist && A.off) { // ghost columns: the owners' column flags travel through the same halo as x
    DevBuf<double> fl((size_t)A.ncols + 2);
    const double *gf = nullptr;
    if (do_cols) {
      { LaunchScope ls(c, "setup"); k_flag_to_double<<<(A.ncols + 255) / 256, 256, 0, c->stream>>>(A.ncols, colflag.p, fl.p); }
      A.halo->begin(fl.p, A.halo_dof);
      A.halo->end();
      gf = A.halo->ghost.p;
    }
    if (A.off->nrows > 0) {
      LaunchScope ls(c, "setup");
      k_zero_offdiag<<<(A.off->nrows + 127) / 128, 128, 0, c->stream>>>(A.off->nrows, A.off->rowptr.p, A.off->col.p, A.off->val.p, A.off_rows.p,
                                                                         do_rows ? rowflag.p : nullptr, gf);
      check_launch("k_zero_offdiag");
    }
    c->sync();
  }
  {
    LaunchScope ls(c, "setup");
    int grid = (A.nrows + 255) / 256;
    k_zero_rows_cols<<<grid, 256, 0, c->stream>>>(A.nrows, A.rowptr.p, A.col.p, A.val.p, do_rows ? rowflag.p : nullptr,
                                                   do_cols ? colflag.p : nullptr, diag, set_diag ? 1 : 0);
    check_launch("k_zero_rows_cols");
  }
  c->sync(); // temporaries are freed on return
}

} // namespace b200sp
