// kernels_spmv.cu -- CSR SpMV for A, B, B^T, C, P, R (replaces MatMult_SeqAIJ / MatMult_MPIAIJ / MatMultAdd reached
// from KSPSolve and PCApply; reference call site src/SaddlePointProblem.c:70).
//
// Algorithmic bytes per launch: 12*nnz + 4*(rows+1) + 8*rows + 8*cols.
// Kernel choice from the row-length histogram (Csr::plan):
//   SPMV_TMA     short rows (default): TMA-staged thread-per-row kernel, kernels_spmv_tma.cu
//   SPMV_STREAM  short rows, no TMA: a warp owns 32 consecutive rows = ONE contiguous segment of val/col.  Phase 1
//                streams that segment with coalesced 128-bit loads, gathers x through L1, parks the products in
//                the warp's shared tile.  Phase 2: lane l sums row l's products in CSR order.
//   SPMV_VECTOR  medium rows: LPR lanes per row (2..32), shuffle-tree reduction.
//   SPMV_BLOCK   very long rows (dense constraint rows): one CTA per row.
// Both short-row kernels are bit-identical to the sequential MatMult_SeqAIJ loop (product rounded, then added).
//
// Columns >= n_owned refer to GHOST values (row-partitioned matrices): the kernels read them from the halo buffer,
// so a distributed MatMult is ONE kernel after the halo exchange, with the same fused epilogues as on one rank:
//   y = beta_z*z + alpha*(A x)                              MatMult / MatMultAdd / residual b - A x
//   y = ca*pm1 + cb*pk + cc*(dinv .* (b - A pk))            one Chebyshev/Jacobi smoothing sweep in a single pass
#include "dev.cuh"
#include "dist.h"
#include <cstdlib>
#include <type_traits>

namespace b200sp {

namespace {

constexpr int STREAM_WARPS = 8;            // 256 threads per CTA
constexpr int STREAM_MAX_GROUP_NNZ = 1152; // per-warp tile limit (9 KB): up to 36 nnz/row on average

__global__ void __launch_bounds__(256) k_spmv_stream(int nrows, const int *__restrict__ rowptr, const int *__restrict__ col,
                                                     const double *__restrict__ val, XSrc xs, double *y, SpmvEpi epi, int tile_elems) {
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
#pragma unroll 4
    for (int base = s_al + 2 * lane; base < e; base += 64) {
      const double2 v = ld_stream_f64x2(val + base);
      const int2 c = ld_stream_s32x2(col + base);
      double2 p;
      p.x = v.x * xs.load(c.x);
      p.y = v.y * xs.load(c.y);
      *reinterpret_cast<double2 *>(prod + (base - s_al)) = p;
    }
    __syncwarp();
    if (r < nrows) {
      double sum = 0.0;
      for (int k = rs - s_al; k < re - s_al; ++k) sum += prod[k];
      y[r] = epi.apply(sum, r);
    }
    __syncwarp();
  }
}

template <int LPR>
__global__ void __launch_bounds__(256) k_spmv_vector(int nrows, const int *__restrict__ rowptr, const int *__restrict__ col,
                                                     const double *__restrict__ val, XSrc xs, double *y, SpmvEpi epi) {
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
      for (int k = rs + sub; k < re; k += LPR) sum += ld_stream_f64(val + k) * xs.load(ld_stream_s32(col + k));
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(FULL, sum, o);
    if (r < nrows && sub == 0) y[r] = epi.apply(sum, r);
  }
}

// One WARP per block row (node) of a matrix with dense BR x BC node blocks and long rows (3-D: 27 blocks, 81 nonzeros per
// row of A): lane b owns block b, loads its BC entries of x ONCE for the BR rows of the node and one block-column id
// instead of BR*BC column ids (8 + 4/(BR*BC) bytes per nonzero instead of 12), keeps BR*BC + BC independent loads in
// flight, and the BR row sums are reduced with a fixed shuffle tree.
template <int BR, int BC>
__global__ void __launch_bounds__(256) k_spmv_nodeblk(int nbrows, const int *__restrict__ rowptr, const int *__restrict__ bptr, const int *__restrict__ bcol,
                                                      const double *__restrict__ val, XSrc xs, double *y, SpmvEpi epi) {
  const int lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
  for (int I = blockIdx.x * wpc + (threadIdx.x >> 5); I < nbrows; I += gridDim.x * wpc) {
    const int b0 = bptr[I], nb = bptr[I + 1] - b0;
    int rs[BR];
    double sum[BR];
#pragma unroll
    for (int rr = 0; rr < BR; ++rr) { rs[rr] = rowptr[I * BR + rr]; sum[rr] = 0.0; }
    for (int b = lane; b < nb; b += 32) {
      const int c0 = ld_stream_s32(bcol + b0 + b) * BC;
      double xv[BC], av[BR][BC];
#pragma unroll
      for (int cc = 0; cc < BC; ++cc) xv[cc] = xs.load(c0 + cc);
#pragma unroll
      for (int rr = 0; rr < BR; ++rr)
#pragma unroll
        for (int cc = 0; cc < BC; ++cc) av[rr][cc] = ld_stream_f64(val + rs[rr] + b * BC + cc);
#pragma unroll
      for (int rr = 0; rr < BR; ++rr)
#pragma unroll
        for (int cc = 0; cc < BC; ++cc) sum[rr] += av[rr][cc] * xv[cc];
    }
#pragma unroll
    for (int rr = 0; rr < BR; ++rr) sum[rr] = warp_sum(sum[rr]);
    if (lane < BR) {
      double s = sum[0];
#pragma unroll
      for (int rr = 1; rr < BR; ++rr)
        if (lane == rr) s = sum[rr];
      y[I * BR + lane] = epi.apply(s, I * BR + lane);
    }
  }
}

__global__ void __launch_bounds__(256) k_spmv_block(int nrows, const int *__restrict__ rowptr, const int *__restrict__ col,
                                                    const double *__restrict__ val, XSrc xs, double *y, SpmvEpi epi) {
  __shared__ double s_w[8];
  for (int r = blockIdx.x; r < nrows; r += gridDim.x) {
    const int rs = rowptr[r], re = rowptr[r + 1];
    double sum = 0.0;
    for (int k = rs + threadIdx.x; k < re; k += 256) sum += ld_stream_f64(val + k) * xs.load(ld_stream_s32(col + k));
    sum = warp_sum(sum);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += s_w[w];
      y[r] = epi.apply(t, r);
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
    for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) // rows are short; ghost columns make the row unsorted on >1 rank
      if (col[k] == r) { v = val[k]; break; }
    d[r] = v;
  }
}

// MatZeroRowsColumns / MatZeroRows / zero columns: one thread per row; owned columns flagged in a byte map, ghost
// columns (>= n_owned) flagged in the halo buffer that carried the owners' flags
__global__ void __launch_bounds__(256) k_zero_rows_cols(int nrows, const int *__restrict__ rowptr, const int *__restrict__ col, double *val,
                                                        const unsigned char *__restrict__ rowflag, const unsigned char *__restrict__ colflag,
                                                        const double *__restrict__ ghostflag, int n_owned, double diag, int set_diag) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x) {
    const bool rz = rowflag && rowflag[r];
    for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) {
      const int c = col[k];
      if (rz) val[k] = (set_diag && c == r) ? diag : 0.0;
      else if (c < n_owned ? (colflag && colflag[c]) : (ghostflag && ghostflag[c - n_owned] != 0.0)) val[k] = 0.0;
    }
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

static bool csr_spmv_launch(const Csr &A, const double *x, double *y, SpmvEpi &epi, bool reuse_halo);
void csr_spmv_epi(const Csr &A, const double *x, double *y, const SpmvEpi &epi_in, bool reuse_halo, Halo *push_to, int push_dof) {
  Ctx *c = A.ctx;
  SpmvEpi epi = epi_in;
  // the produced vector is multiplied next by a matrix with the halo *push_to: push its boundary values now.  Only for
  // peer-to-peer halos; fused into the TMA kernels, otherwise the ordinary push kernel right after this SpMV.
  static const bool no_fused_push = getenv("B200SP_NO_FUSED_PUSH") && atoi(getenv("B200SP_NO_FUSED_PUSH"));
  if (push_to && !(c->dcomm && push_to->p2p && A.nrows > 0) ) push_to = nullptr;
  if (push_to && no_fused_push) push_to = nullptr;
  if (push_to && A.kernel == SPMV_TMA) epi.push = push_to->push_out(push_dof);
  const bool fused = csr_spmv_launch(A, x, y, epi, reuse_halo);
  if (push_to) {
    if (!fused) push_to->begin(y, push_dof); // the kernel that ran cannot push: ordinary push kernel, still ahead of the consumer
    push_to->pushed_vec = y;
  }
}

// returns true when the kernel that ran also pushed the produced vector (epi.push)
static bool csr_spmv_launch(const Csr &A, const double *x, double *y, SpmvEpi &epi, bool reuse_halo) {
  Ctx *c = A.ctx;
  XSrc xs{x, nullptr, 0x7fffffff};
  bool pending_wait = false;
  if (A.halo && c->dcomm && A.halo->p2p) {
    // peer-to-peer halo: push (unless the ghosts of this x are already there); the TMA kernel waits for the
    // neighbours itself, other kernels get the separate wait kernel
    if (!reuse_halo) A.halo->begin(x, A.halo_dof);
    xs.ghost = A.halo->ghost.p;
    xs.n_owned = A.ncols;
    xs.seq = A.halo->seq.p;
    xs.ghost_stride = A.halo->ghost_stride;
    if (!reuse_halo) {
      if (A.kernel == SPMV_TMA) { xs.wait_flags = A.halo->flags.p; xs.wait_nmsg = A.halo->n_msgs; xs.wait_err = c->d_err; pending_wait = true; }
      else A.halo->end();
    }
  } else if (A.halo && c->dcomm) {
    // MatMult_MPIAIJ.  The ghost values of x travel on the halo stream while the tiles that have no ghost column
    // (the interior: ~90% of the rows) are multiplied; the boundary tiles follow once the halo has arrived.
    epi.push = PushOut(); // send/recv halos: no fused push (push_to is peer-to-peer only, so it is null here anyway)
    if (!reuse_halo) A.halo->begin(x, A.halo_dof);
    // Measured (profiles/r01_scaling_notes.md): with the persistent TMA grid holding every SM the NCCL kernel cannot
    // start until the interior kernel drains, so the two-launch split was slightly SLOWER (20.3 vs 19.9 ms at 8 GPUs).
    // It stays available behind B200SP_SPMV_OVERLAP=1; the default is one kernel after the halo has arrived.
    static const bool want_split = getenv("B200SP_SPMV_OVERLAP") && atoi(getenv("B200SP_SPMV_OVERLAP")) != 0;
    const bool split = want_split && A.kernel == SPMV_TMA && A.tiles_interior.p && spmv_tma_tile_rows() == TMA_TILE_ROWS && A.nrows > 0;
    if (split && A.n_tiles_interior > 0) {
      LaunchScope ls(c, A.tag.c_str());
      csr_spmv_tma(A, xs, y, epi, A.tiles_interior.p, A.n_tiles_interior);
    }
    A.halo->end();
    xs.ghost = A.halo->ghost.p;
    xs.n_owned = A.ncols;
    if (A.halo->p2p) { xs.seq = A.halo->seq.p; xs.ghost_stride = A.halo->ghost_stride; }
    if (split) {
      if (A.n_tiles_boundary > 0) {
        LaunchScope ls(c, A.tag.c_str());
        csr_spmv_tma(A, xs, y, epi, A.tiles_boundary.p, A.n_tiles_boundary);
      }
      return false;
    }
  }
  if (A.nrows <= 0) return false;
  // measurement pass: one profile class per matrix AND epilogue variant (they move different numbers of vectors)
  std::string cls_variant;
  if (c->profile) cls_variant = A.tag + (epi.cheb ? "|cheb" : epi.z ? "|axpby" : "|plain");
  LaunchScope ls(c, c->profile ? cls_variant.c_str() : A.tag.c_str());
  if (A.kernel == SPMV_TMA && csr_spmv_tma(A, xs, y, epi)) return epi.push.grp != nullptr;
  epi.push = PushOut(); // the kernels below do not push
  if (pending_wait) { A.halo->end(); xs.wait_flags = nullptr; } // the TMA kernel declined: wait with the separate kernel
  // long rows with repeating values (the 3-D operators: 81 entries per row): scalar-row tile dictionaries
  if ((A.kernel == SPMV_NODE || A.kernel == SPMV_VECTOR) && A.max_row_nnz <= 96 && csr_spmv_pd(A, xs, y, epi)) return false;
  if (A.kernel == SPMV_STREAM || A.kernel == SPMV_TMA) {
    int tile = (A.max_group_nnz + 1) & ~1;
    size_t smem = (size_t)tile * sizeof(double) * STREAM_WARPS;
    if (!(c->attr_mask & 1u)) {
      B2_CUDA(cudaFuncSetAttribute(k_spmv_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, STREAM_MAX_GROUP_NNZ * 8 * STREAM_WARPS + 64));
      c->attr_mask |= 1u;
    }
    int per_sm = (int)((220 * 1024) / (smem + 1024));
    if (per_sm > 8) per_sm = 8;
    if (per_sm < 1) per_sm = 1;
    int ngroups = (A.nrows + 31) / 32;
    int grid = (ngroups + STREAM_WARPS - 1) / STREAM_WARPS;
    if (grid > c->num_sms * per_sm) grid = c->num_sms * per_sm;
    k_spmv_stream<<<grid, 256, smem, c->stream>>>(A.nrows, A.rowptr.p, A.col.p, A.val.p, xs, y, epi, tile);
    check_launch("k_spmv_stream");
  } else if (A.kernel == SPMV_NODE && A.bcol.p) {
    const int nb = A.nrows / A.blk_r;
    int grid = (nb + 7) / 8;
    if (grid > c->num_sms * 8) grid = c->num_sms * 8;
    if (A.blk_r == 3 && A.blk_c == 3) k_spmv_nodeblk<3, 3><<<grid, 256, 0, c->stream>>>(nb, A.rowptr.p, A.bptr.p, A.bcol.p, A.val.p, xs, y, epi);
    else if (A.blk_r == 3 && A.blk_c == 1) k_spmv_nodeblk<3, 1><<<grid, 256, 0, c->stream>>>(nb, A.rowptr.p, A.bptr.p, A.bcol.p, A.val.p, xs, y, epi);
    else if (A.blk_r == 1 && A.blk_c == 3) k_spmv_nodeblk<1, 3><<<grid, 256, 0, c->stream>>>(nb, A.rowptr.p, A.bptr.p, A.bcol.p, A.val.p, xs, y, epi);
    else throw Error(B200SP_ERR_UNSUPPORTED, "node-block SpMV: unsupported block shape");
    check_launch("k_spmv_nodeblk");
  } else if (A.kernel == SPMV_BLOCK) {
    int grid = A.nrows < c->num_sms * 8 ? A.nrows : c->num_sms * 8;
    k_spmv_block<<<grid, 256, 0, c->stream>>>(A.nrows, A.rowptr.p, A.col.p, A.val.p, xs, y, epi);
    check_launch("k_spmv_block");
  } else {
    auto launch = [&](auto lpr_tag) {
      constexpr int LPR = decltype(lpr_tag)::value;
      int rows_per_cta = 256 / LPR;
      int grid = (A.nrows + rows_per_cta - 1) / rows_per_cta;
      if (grid > c->num_sms * 8) grid = c->num_sms * 8;
      k_spmv_vector<LPR><<<grid, 256, 0, c->stream>>>(A.nrows, A.rowptr.p, A.col.p, A.val.p, xs, y, epi);
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
  return false;
}

void csr_spmv(const Csr &A, const double *x, double *y, double alpha, const double *z, double beta_z, bool reuse_halo, Halo *push_to, int push_dof) {
  SpmvEpi e;
  e.alpha = alpha; e.z = z; e.beta_z = beta_z;
  csr_spmv_epi(A, x, y, e, reuse_halo, push_to, push_dof);
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
  DevBuf<double> fl;
  const double *ghostflag = nullptr;
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
    { LaunchScope ls(c, "setup"); k_mark<<<g, 256, 0, c->stream>>>(n, d_idx.p, colflag.p); }
    if (dist) { // the owners' column flags travel through the same halo as x
      fl.alloc((size_t)A.ncols + 2);
      { LaunchScope ls(c, "setup"); k_flag_to_double<<<(A.ncols + 255) / 256, 256, 0, c->stream>>>(A.ncols, colflag.p, fl.p); }
      A.halo->begin(fl.p, A.halo_dof);
      A.halo->end();
      ghostflag = A.halo->ghost_now();
    }
  }
  {
    LaunchScope ls(c, "setup");
    int grid = (A.nrows + 255) / 256;
    k_zero_rows_cols<<<grid, 256, 0, c->stream>>>(A.nrows, A.rowptr.p, A.col.p, A.val.p, do_rows ? rowflag.p : nullptr,
                                                   do_cols ? colflag.p : nullptr, ghostflag, A.ncols, diag, set_diag ? 1 : 0);
    check_launch("k_zero_rows_cols");
  }
  c->sync(); // temporaries are freed on return
  csr_drop_value_dict(A); // values changed
  A.state++;
}

} // namespace b200sp
