// kernels_spmv_tma.cu -- SPMV_TMA: CSR SpMV whose matrix stream is moved by the TMA engine.
//
// A persistent CTA walks row tiles of R = blockDim.x rows.  The tile's val/col segment (one contiguous
// range of the CSR arrays) is copied into shared memory by two cp.async.bulk (UBLKCP) requests issued by
// one thread and tracked by an mbarrier (complete_tx); S stages are in flight, so the HBM stream never
// waits for the arithmetic and costs no registers or issue slots.  Consumers: one thread per row walks its
// row in CSR order out of shared memory, gathers x through L1/L2 and accumulates product-then-add, i.e.
// bit-identical to the sequential MatMult_SeqAIJ loop (same result as SPMV_STREAM).
//   shared memory per stage = 12 B x (max tile nnz + pad); A block (18 nnz/row, R = 512): 110.7 KB x 2 stages.
//   algorithmic bytes per launch: 12 nnz + 4 (rows+1) + 8 rows + 8 cols  (SURVEY 8d).
#include "dev.cuh"
#include <algorithm>
#include <cstdlib>

namespace b200sp {

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
  } while (!ok);
}

constexpr int TMA_MAX_STAGES = 4;

// dynamic shared layout: [S stages][ val: cap doubles | col: cap ints ], then S mbarriers
template <int UNROLL>
__global__ void k_spmv_tma(int nrows, int ntiles, const int *__restrict__ tile_list, const int *__restrict__ rowptr, const int *__restrict__ col,
                           const double *__restrict__ val, XSrc xs, double *y, SpmvEpi epi, int cap, int stages) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  const int R = blockDim.x;
  const size_t stage_bytes = (size_t)cap * 12;
  uint64_t *full = reinterpret_cast<uint64_t *>(s_raw + stage_bytes * stages);
  const int tid = threadIdx.x;

  // ntiles counts the entries of tile_list when one is given (interior / boundary subsets of a row-partitioned
  // matrix, so the interior can run while the halo is in flight), otherwise all tiles 0..ntiles-1
  auto issue = [&](int idx, int stage) { // one thread
    const int tile = tile_list ? tile_list[idx] : idx;
    const int r0 = tile * R;
    const int r1 = min(r0 + R, nrows);
    const int s0 = rowptr[r0] & ~3;
    const int e0 = rowptr[r1];
    const int cnt = (e0 - s0 + 3) & ~3;
    unsigned char *base = s_raw + stage_bytes * stage;
    mbar_expect_tx(&full[stage], (unsigned)cnt * 12u);
    if (cnt > 0) {
      bulk_g2s(base, val + s0, (unsigned)cnt * 8u, &full[stage]);
      bulk_g2s(base + (size_t)cap * 8, col + s0, (unsigned)cnt * 4u, &full[stage]);
    }
  };

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      const int t = blockIdx.x + s * gridDim.x;
      if (t < ntiles) issue(t, s);
    }
  }
  if (xs.wait_flags) { // peer-to-peer halo: the first tiles are already in flight while we wait for the neighbours
    if (tid < xs.wait_nmsg) {
      const unsigned long long want = *xs.seq;
      volatile const unsigned long long *f = xs.wait_flags + tid;
      for (long long spin = 0; *f < want; ++spin) {
        __nanosleep(100);
        if (spin > 20000000LL) { *xs.wait_err = 400 + tid; __threadfence_system(); break; }
      }
      __threadfence_system();
    }
    __syncthreads();
  }
  int it = 0;
  for (int idx = blockIdx.x; idx < ntiles; idx += gridDim.x, ++it) {
    const int tile = tile_list ? tile_list[idx] : idx;
    const int stage = it % stages;
    const unsigned parity = (unsigned)(it / stages) & 1u;
    const int r = tile * R + tid;
    // row bounds are fetched before waiting on the tile so their latency overlaps the TMA
    const int s_al = rowptr[tile * R] & ~3;
    const int rs = rowptr[r < nrows ? r : nrows];
    const int re = rowptr[r + 1 < nrows ? r + 1 : nrows];
    mbar_wait(&full[stage], parity);
    const double *sv = reinterpret_cast<const double *>(s_raw + stage_bytes * stage);
    const int *sc = reinterpret_cast<const int *>(s_raw + stage_bytes * stage + (size_t)cap * 8);
    if (r < nrows) {
      double sum = 0.0;
      int k = rs - s_al;
      const int ke = re - s_al;
      for (; k + UNROLL <= ke; k += UNROLL) {
        double xv[UNROLL], av[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) { av[u] = sv[k + u]; xv[u] = xs.load(sc[k + u]); }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) sum += av[u] * xv[u];
      }
      for (; k < ke; ++k) sum += sv[k] * xs.load(sc[k]);
      y[r] = epi.apply(sum, r);
    }
    __syncthreads(); // every consumer is done with this stage
    if (tid == 0) {
      const int t = idx + stages * gridDim.x;
      if (t < ntiles) issue(t, stage);
    }
  }
}

// ---- block-compressed column index (BCSR-style indices, CSR values) ---------------------------------------------
// The DMDA matrices have dense BR x BC node blocks (A: 2x2, B: 1x2, B^T: 2x1): the BR rows of a node share their
// columns and the columns come in aligned runs of BC.  The values stay in CSR order (so the per-row summation order,
// and therefore every bit of the result, is unchanged) but the kernel streams ONE block-column index per BR x BC
// entries instead of one column index per entry: 8 + 4/(BR*BC) bytes per nonzero instead of 12 (A: 9 B/nnz, -25%).
// The CSR col array is kept for MatView/export and for the other kernels.
template <int BR, int BC, int UNROLL>
__global__ void k_spmv_tma_blk(int nrows, int ntiles, const int *__restrict__ tile_list, const int *__restrict__ rowptr, const int *__restrict__ bptr,
                               const int *__restrict__ bcol, const double *__restrict__ val, XSrc xs, double *y, SpmvEpi epi, int cap, int capb, int stages) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  const int R = blockDim.x;
  const size_t stage_bytes = (size_t)cap * 8 + (size_t)capb * 4;
  uint64_t *full = reinterpret_cast<uint64_t *>(s_raw + stage_bytes * stages);
  const int tid = threadIdx.x;
  auto issue = [&](int idx, int stage) { // one thread
    const int tile = tile_list ? tile_list[idx] : idx;
    const int r0 = tile * R;
    const int r1 = min(r0 + R, nrows);
    const int s0 = rowptr[r0] & ~3;
    const int cnt = (rowptr[r1] - s0 + 3) & ~3;
    const int b0 = bptr[r0 / BR] & ~3;
    const int bcnt = (bptr[r1 / BR] - b0 + 3) & ~3;
    unsigned char *base = s_raw + stage_bytes * stage;
    mbar_expect_tx(&full[stage], (unsigned)cnt * 8u + (unsigned)bcnt * 4u);
    if (cnt > 0) bulk_g2s(base, val + s0, (unsigned)cnt * 8u, &full[stage]);
    if (bcnt > 0) bulk_g2s(base + (size_t)cap * 8, bcol + b0, (unsigned)bcnt * 4u, &full[stage]);
  };
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      const int t = blockIdx.x + s * gridDim.x;
      if (t < ntiles) issue(t, s);
    }
  }
  if (xs.wait_flags) {
    if (tid < xs.wait_nmsg) {
      const unsigned long long want = *xs.seq;
      volatile const unsigned long long *f = xs.wait_flags + tid;
      for (long long spin = 0; *f < want; ++spin) {
        __nanosleep(100);
        if (spin > 20000000LL) { *xs.wait_err = 400 + tid; __threadfence_system(); break; }
      }
      __threadfence_system();
    }
    __syncthreads();
  }
  int it = 0;
  for (int idx = blockIdx.x; idx < ntiles; idx += gridDim.x, ++it) {
    const int tile = tile_list ? tile_list[idx] : idx;
    const int stage = it % stages;
    const unsigned parity = (unsigned)(it / stages) & 1u;
    const int r = tile * R + tid;
    const int s_al = rowptr[tile * R] & ~3;
    const int b_al = bptr[tile * R / BR] & ~3;
    const int rs = rowptr[r < nrows ? r : nrows];
    const int re = rowptr[r + 1 < nrows ? r + 1 : nrows];
    const int bs = bptr[(r < nrows ? r : nrows) / BR];
    mbar_wait(&full[stage], parity);
    const double *sv = reinterpret_cast<const double *>(s_raw + stage_bytes * stage) + (rs - s_al);
    const int *sb = reinterpret_cast<const int *>(s_raw + stage_bytes * stage + (size_t)cap * 8) + (bs - b_al);
    if (r < nrows) {
      double sum = 0.0;
      const int len = re - rs;
      int k = 0;
      for (; k + UNROLL <= len; k += UNROLL) { // UNROLL is a multiple of BC: whole blocks per step
        double xv[UNROLL], av[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) { av[u] = sv[k + u]; xv[u] = xs.load(sb[(k + u) / BC] * BC + (u % BC)); }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) sum += av[u] * xv[u];
      }
      for (; k < len; ++k) sum += sv[k] * xs.load(sb[k / BC] * BC + (k % BC));
      y[r] = epi.apply(sum, r);
    }
    __syncthreads();
    if (tid == 0) {
      const int t = idx + stages * gridDim.x;
      if (t < ntiles) issue(t, stage);
    }
  }
}

// verify the block structure of block row I and count its blocks
template <int BR, int BC>
__global__ void __launch_bounds__(256) k_blk_check(int nbrows, const int *__restrict__ rowptr, const int *__restrict__ col, int *bcnt, int *fail) {
  for (int I = blockIdx.x * blockDim.x + threadIdx.x; I < nbrows; I += gridDim.x * blockDim.x) {
    const int r0 = I * BR, rs0 = rowptr[r0], L = rowptr[r0 + 1] - rs0;
    bool ok = L % BC == 0;
    for (int rr = 1; rr < BR && ok; ++rr) ok = rowptr[r0 + rr + 1] - rowptr[r0 + rr] == L;
    for (int b = 0; b < L / BC && ok; ++b) {
      const int c0 = col[rs0 + b * BC];
      ok = c0 % BC == 0;
      for (int rr = 0; rr < BR && ok; ++rr)
        for (int cc = 0; cc < BC && ok; ++cc) ok = col[rowptr[r0 + rr] + b * BC + cc] == c0 + cc;
    }
    if (!ok) *fail = 1;
    bcnt[I] = ok ? L / BC : 0;
  }
}
template <int BR, int BC>
__global__ void __launch_bounds__(256) k_blk_fill(int nbrows, const int *__restrict__ rowptr, const int *__restrict__ col, const int *__restrict__ bptr, int *bcol) {
  for (int I = blockIdx.x * blockDim.x + threadIdx.x; I < nbrows; I += gridDim.x * blockDim.x) {
    const int rs0 = rowptr[I * BR], nb = bptr[I + 1] - bptr[I];
    for (int b = 0; b < nb; ++b) bcol[bptr[I] + b] = col[rs0 + b * BC] / BC;
  }
}

template <int BR, int BC>
bool build_block_index(Csr &A) {
  Ctx *c = A.ctx;
  if (A.nrows % BR != 0 || A.nrows == 0) return false;
  const int nb = A.nrows / BR;
  DevBuf<int> cnt((size_t)nb + 1), fail(1);
  fail.zero(c->stream);
  const int grid = std::max(1, std::min((nb + 255) / 256, c->num_sms * 16));
  { LaunchScope ls(c, "setup"); k_blk_check<BR, BC><<<grid, 256, 0, c->stream>>>(nb, A.rowptr.p, A.col.p, cnt.p, fail.p); check_launch("k_blk_check"); }
  int h_fail = 0;
  B2_CUDA(cudaMemcpyAsync(&h_fail, fail.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  c->sync();
  if (h_fail) return false;
  A.bptr.alloc((size_t)nb + 1);
  int total = 0;
  exclusive_scan_i32(c, cnt.p, A.bptr.p, nb, &total);
  A.bcol.alloc((size_t)total + CSR_PAD);
  A.bcol.zero(c->stream);
  { LaunchScope ls(c, "setup"); k_blk_fill<BR, BC><<<grid, 256, 0, c->stream>>>(nb, A.rowptr.p, A.col.p, A.bptr.p, A.bcol.p); check_launch("k_blk_fill"); }
  c->sync();
  A.blk_r = BR; A.blk_c = BC;
  return true;
}

} // namespace

bool csr_try_block_index(Csr &A, int br, int bc) {
  static const bool off = getenv("B200SP_NO_BLOCK_INDEX") && atoi(getenv("B200SP_NO_BLOCK_INDEX"));
  if (off || A.kernel != SPMV_TMA) return false;
  if (br == 2 && bc == 2) return build_block_index<2, 2>(A);
  if (br == 1 && bc == 2) return build_block_index<1, 2>(A);
  if (br == 2 && bc == 1) return build_block_index<2, 1>(A);
  return false;
}

// returns false when the matrix does not fit the shared-memory tiling (caller falls back)
int spmv_tma_tile_rows() {
  static const int env_R = getenv("B200SP_TMA_R") ? atoi(getenv("B200SP_TMA_R")) : 0;
  return env_R ? env_R : TMA_TILE_ROWS;
}

bool csr_spmv_tma(const Csr &A, const XSrc &xs, double *y, const SpmvEpi &epi, const int *tile_list, int nlist) {
  Ctx *c = A.ctx;
  const size_t budget = 225 * 1024;
  int R = 0, stages = 0, cap = 0;
  static const int env_R = getenv("B200SP_TMA_R") ? atoi(getenv("B200SP_TMA_R")) : 0;           // tuning overrides
  static const int env_S = getenv("B200SP_TMA_STAGES") ? atoi(getenv("B200SP_TMA_STAGES")) : 0;
  static const int env_U = getenv("B200SP_TMA_UNROLL") ? atoi(getenv("B200SP_TMA_UNROLL")) : 0;
  // Measured on B200 (profiles/r01_tma_tile_sweep.txt): small tiles with exactly two stages and as many
  // co-resident CTAs as shared memory allows beat large tiles / deeper pipelines for every block
  // (A 18 nnz/row: R=128,S=2 -> 97% of the measured HBM peak; R=512,S=2 -> 93%; any S>=3 -> <= 62%).
  for (int r : {128, 256, 512}) {
    if (env_R && r != env_R) continue;
    const int capr = (((r / 32) * A.max_group_nnz + 8) + 3) & ~3;
    const size_t sb = (size_t)capr * 12;
    int s = (int)((budget - 64) / sb);
    if (s >= 2) { R = r; stages = 2; cap = capr; if (env_S && env_S <= s && env_S <= TMA_MAX_STAGES) stages = env_S; break; }
    if (!env_R) break; // rows too long for a 128-row tile: the vector kernel is the right tool
  }
  if (!R) return false;
  const size_t smem = (size_t)cap * 12 * stages + 8 * TMA_MAX_STAGES;
  static bool attr_set = false;
  if (!attr_set) {
    B2_CUDA(cudaFuncSetAttribute(k_spmv_tma<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    B2_CUDA(cudaFuncSetAttribute(k_spmv_tma<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  if (tile_list && R != TMA_TILE_ROWS) return false; // the lists were built for TMA_TILE_ROWS-row tiles
  const int ntiles = tile_list ? nlist : (A.nrows + R - 1) / R;
  if (ntiles <= 0) return true;
  int per_sm = (int)(budget / smem);
  if (per_sm < 1) per_sm = 1;
  if (per_sm * R > 2048) per_sm = 2048 / R;
  int grid = ntiles < c->num_sms * per_sm ? ntiles : c->num_sms * per_sm;
  const double mean = A.nrows ? (double)A.nnz / A.nrows : 0.0;
  (void)mean;
  if (A.bcol.p && R == TMA_TILE_ROWS) { // block-compressed column index: 8 + 4/(BR*BC) bytes per nonzero
    const int capb = ((cap / (A.blk_r * A.blk_c) + 8) + 3) & ~3;
    const size_t smem_b = ((size_t)cap * 8 + (size_t)capb * 4) * stages + 8 * TMA_MAX_STAGES;
    static bool attr_b = false;
    if (!attr_b) {
      B2_CUDA(cudaFuncSetAttribute(k_spmv_tma_blk<2, 2, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      B2_CUDA(cudaFuncSetAttribute(k_spmv_tma_blk<1, 2, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      B2_CUDA(cudaFuncSetAttribute(k_spmv_tma_blk<2, 1, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      attr_b = true;
    }
    int psm = (int)(budget / smem_b);
    if (psm < 1) psm = 1;
    if (psm * R > 2048) psm = 2048 / R;
    const int gridb = ntiles < c->num_sms * psm ? ntiles : c->num_sms * psm;
    if (A.blk_r == 2 && A.blk_c == 2)
      k_spmv_tma_blk<2, 2, 6><<<gridb, R, smem_b, c->stream>>>(A.nrows, ntiles, tile_list, A.rowptr.p, A.bptr.p, A.bcol.p, A.val.p, xs, y, epi, cap, capb, stages);
    else if (A.blk_r == 1 && A.blk_c == 2)
      k_spmv_tma_blk<1, 2, 6><<<gridb, R, smem_b, c->stream>>>(A.nrows, ntiles, tile_list, A.rowptr.p, A.bptr.p, A.bcol.p, A.val.p, xs, y, epi, cap, capb, stages);
    else
      k_spmv_tma_blk<2, 1, 6><<<gridb, R, smem_b, c->stream>>>(A.nrows, ntiles, tile_list, A.rowptr.p, A.bptr.p, A.bcol.p, A.val.p, xs, y, epi, cap, capb, stages);
    check_launch("k_spmv_tma_blk");
    return true;
  }
  if (env_U ? env_U == 6 : true)
    k_spmv_tma<6><<<grid, R, smem, c->stream>>>(A.nrows, ntiles, tile_list, A.rowptr.p, A.col.p, A.val.p, xs, y, epi, cap, stages);
  else
    k_spmv_tma<3><<<grid, R, smem, c->stream>>>(A.nrows, ntiles, tile_list, A.rowptr.p, A.col.p, A.val.p, xs, y, epi, cap, stages);
  check_launch("k_spmv_tma");
  return true;
}

} // namespace b200sp
