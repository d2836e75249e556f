// kernels_spmv_tma.cu -- SPMV_TMA: CSR SpMV whose matrix stream is moved by the TMA engine.
//
// A persistent CTA walks row tiles of R = blockDim.x rows.  The tile's val/col segment (one contiguous
// range of the CSR arrays) is copied into shared memory by two cp.async.bulk (UBLKCP) requests issued by
// one thread and tracked by an mbarrier (complete_tx); S stages are in flight, so the HBM stream never
// waits for the arithmetic and costs no registers or issue slots.  Consumers: one thread per row walks its
// row in CSR order out of shared memory, gathers x through L1/L2 and accumulates product-then-add, i.e.
// bit-identical to the sequential MatMult_SeqAIJ loop (same result as SPMV_STREAM).
//   shared memory per stage = 12 B x (max tile nnz + pad); A block (18 nnz/row, R = 128): 27.8 KB x 2 stages, 4 CTAs/SM.
//   algorithmic bytes per launch: 12 nnz + 4 (rows+1) + 8 rows + 8 cols  (SURVEY 8d).
//
// Three kernels share that pipeline and differ in WHAT is streamed (all give bit-identical results):
//   k_spmv_tma        values 8 B + column 4 B per nonzero                      (plain CSR; C, Q, unstructured matrices)
//   k_spmv_tma_blk    values 8 B + one block-column id per BR x BC node block   (A 9 B/nnz; when the dictionary declines)
//   k_spmv_tma_dict   16-bit codes into a per-tile dictionary of distinct values + block-column ids, one thread per
//                     node block row (A 3.74 B/nnz; default for A, B^T, B, P, R)
// Row-partitioned matrices: ghost columns are read from the halo buffer (XSrc); with peer-to-peer halos the kernel
// itself waits for the neighbours' flags when it reaches the first tile that has a ghost column (tiles without come
// first, Csr::wait_order) and can push the boundary rows of the vector it produces to the neighbours (SpmvEpi::push).
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

// peer-to-peer halo: one lane per incoming message polls this rank's flag words until the neighbours' pushes of the
// current exchange have landed (bounded; reports through the mapped error word).  Called by the whole CTA.
__device__ __forceinline__ void halo_wait_cta(const XSrc &xs) {
  if ((int)threadIdx.x < xs.wait_nmsg) {
    const unsigned long long want = *xs.seq;
    volatile const unsigned long long *f = xs.wait_flags + threadIdx.x;
    for (long long spin = 0; *f < want; ++spin) {
      __nanosleep(100);
      if (spin > 20000000LL) { *xs.wait_err = 400 + (int)threadIdx.x; __threadfence_system(); break; }
    }
    __threadfence_system();
  }
  __syncthreads();
}

// dynamic shared layout: [S stages][ val: cap doubles | col: cap ints ], then S mbarriers
template <int UNROLL>
__global__ void k_spmv_tma(int nrows, int ntiles, const int *__restrict__ tile_list, const int *__restrict__ rowptr, const int *__restrict__ col,
                           const double *__restrict__ val, XSrc xs, double *y, SpmvEpi epi, int cap, int stages, int n_nowait) {
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
  bool waited = xs.wait_flags == nullptr; // tiles before n_nowait read no ghost column: the wait is deferred until the first one that does
  int it = 0;
  for (int idx = blockIdx.x; idx < ntiles; idx += gridDim.x, ++it) {
    const int tile = tile_list ? tile_list[idx] : idx;
    if (!waited && idx >= n_nowait) { halo_wait_cta(xs); waited = true; }
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
      if (k < ke) { // tail: the remaining (< UNROLL) entries are loaded together under predicates, not one by one
        double xv[UNROLL - 1], av[UNROLL - 1];
#pragma unroll
        for (int u = 0; u < UNROLL - 1; ++u)
          if (k + u < ke) { av[u] = sv[k + u]; xv[u] = xs.load(sc[k + u]); }
#pragma unroll
        for (int u = 0; u < UNROLL - 1; ++u)
          if (k + u < ke) sum += av[u] * xv[u];
      }
      const double v = epi.apply(sum, r);
      y[r] = v;
      if (epi.push.grp) epi.push.row(r, v);
    }
    __syncthreads(); // every consumer is done with this stage
    if (tid == 0) {
      const int t = idx + stages * gridDim.x;
      if (t < ntiles) issue(t, stage);
    }
  }
  if (epi.push.grp) epi.push.finish(); // fused halo push: fence, ticket, last CTA raises the neighbours' flags
}

// ---- block-compressed column index (BCSR-style indices, CSR values) ---------------------------------------------
// The DMDA matrices have dense BR x BC node blocks (A: 2x2, B: 1x2, B^T: 2x1): the BR rows of a node share their
// columns and the columns come in aligned runs of BC.  The values stay in CSR order (so the per-row summation order,
// and therefore every bit of the result, is unchanged) but the kernel streams ONE block-column index per BR x BC
// entries instead of one column index per entry: 8 + 4/(BR*BC) bytes per nonzero instead of 12 (A: 9 B/nnz, -25%).
// The CSR col array is kept for MatView/export and for the other kernels.
template <int BR, int BC, int UNROLL>
__global__ void k_spmv_tma_blk(int nrows, int ntiles, const int *__restrict__ tile_list, const int *__restrict__ rowptr, const int *__restrict__ bptr,
                               const int *__restrict__ bcol, const double *__restrict__ val, XSrc xs, double *y, SpmvEpi epi, int cap, int capb, int stages, int n_nowait) {
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
  bool waited = xs.wait_flags == nullptr; // tiles before n_nowait read no ghost column: the wait is deferred until the first one that does
  int it = 0;
  for (int idx = blockIdx.x; idx < ntiles; idx += gridDim.x, ++it) {
    const int tile = tile_list ? tile_list[idx] : idx;
    if (!waited && idx >= n_nowait) { halo_wait_cta(xs); waited = true; }
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
      if (k < len) { // tail under predicates (k is a multiple of BC here)
        double xv[UNROLL - 1], av[UNROLL - 1];
#pragma unroll
        for (int u = 0; u < UNROLL - 1; ++u)
          if (k + u < len) { av[u] = sv[k + u]; xv[u] = xs.load(sb[(k + u) / BC] * BC + (u % BC)); }
#pragma unroll
        for (int u = 0; u < UNROLL - 1; ++u)
          if (k + u < len) sum += av[u] * xv[u];
      }
      const double v = epi.apply(sum, r);
      y[r] = v;
      if (epi.push.grp) epi.push.row(r, v);
    }
    __syncthreads();
    if (tid == 0) {
      const int t = idx + stages * gridDim.x;
      if (t < ntiles) issue(t, stage);
    }
  }
  if (epi.push.grp) epi.push.finish(); // fused halo push: fence, ticket, last CTA raises the neighbours' flags
}

// ---- tile-local value dictionary (CSR-VI style value indexing, per tile) ------------------------------------------
// Finite-element matrices on structured grids repeat a small set of element-matrix sums: a 128-row tile of the A block
// holds ~2300 values but only ~300 DISTINCT bit patterns.  Each tile therefore stores its distinct values once (8 B
// each, in order of first occurrence) and one 16-bit code per nonzero; the kernel streams dictionary + codes + column
// index (A: ~3.7 B/nnz instead of 9) and looks the value up in shared memory.  The value fed to the multiply is the
// identical double, in the identical CSR order, so the result is bit-for-bit the same as every other SpMV kernel.
// Matrices whose tiles do not compress (unstructured values) keep the plain value stream.
constexpr int DICT_T = 8192;             // hash slots per tile in the build kernel
constexpr int DICT_MAX_TILE_NNZ = 6144;  // larger tiles: no dictionary
constexpr int DICT_MAX_ENTRIES = 2048;   // per tile (8 KB of shared memory per stage at most)
constexpr unsigned long long DICT_EMPTY = ~0ull;

// One thread per BLOCK ROW (node): the BR rows of a node share their block columns, so one x load (16 bytes when
// BC = 2) and one index load serve BR x BC nonzeros -- the kernel is bound by L1/shared-memory wavefronts once the
// stream is this small, not by HBM.  Each row still accumulates its own entries in CSR order.  A tile is blockDim.x
// block rows (BR * 128 rows); rowptr is not needed: row I*BR + rr starts at (bptr[I] * BR + rr * nb) * BC.
// BR = BC = 1 is plain CSR (bptr = rowptr, bcol = col).
template <int BR, int BC, int UB>
__global__ void k_spmv_tma_dict(int nbrows, int ntiles, const int *__restrict__ tile_list, const int *__restrict__ bptr, const int *__restrict__ bcol,
                                const int *__restrict__ dptr, const double *__restrict__ dict, const unsigned short *__restrict__ codes, XSrc xs,
                                double *y, SpmvEpi epi, int capc, int capb, int dcap, int stages, int n_nowait) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  const int NB = blockDim.x;
  const size_t off_codes = (size_t)dcap * 8, off_bcol = off_codes + (size_t)capc * 2; // capc is a multiple of 8
  const size_t stage_bytes = off_bcol + (size_t)capb * 4;
  uint64_t *full = reinterpret_cast<uint64_t *>(s_raw + stage_bytes * stages);
  const int tid = threadIdx.x;
  auto issue = [&](int idx, int stage) { // one thread
    const int tile = tile_list ? tile_list[idx] : idx;
    const int I0 = tile * NB;
    const int I1 = min(I0 + NB, nbrows);
    const int g0 = bptr[I0], g1 = bptr[I1];
    const int s0 = (g0 * (BR * BC)) & ~7;
    const int cnt = (g1 * (BR * BC) - s0 + 7) & ~7;
    const int b0 = g0 & ~3;
    const int bcnt = (g1 - b0 + 3) & ~3;
    const int d0 = dptr[tile], dcnt = dptr[tile + 1] - d0;
    unsigned char *base = s_raw + stage_bytes * stage;
    mbar_expect_tx(&full[stage], (unsigned)dcnt * 8u + (unsigned)cnt * 2u + (unsigned)bcnt * 4u);
    if (dcnt > 0) bulk_g2s(base, dict + d0, (unsigned)dcnt * 8u, &full[stage]);
    if (cnt > 0) bulk_g2s(base + off_codes, codes + s0, (unsigned)cnt * 2u, &full[stage]);
    if (bcnt > 0) bulk_g2s(base + off_bcol, bcol + b0, (unsigned)bcnt * 4u, &full[stage]);
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
  bool waited = xs.wait_flags == nullptr; // tiles before n_nowait read no ghost column: the wait is deferred until the first one that does
  // values of block b of this thread's rows and the x entries they multiply
  auto fetch = [&](const double *sd, const unsigned short *sc, const int *sb, int L, int b, double (&av)[BR][BC], double (&xv)[BC]) {
    const int c0 = sb[b] * BC;
    if (BC == 2) {
      const double2 x2 = xs.load2(c0);
      xv[0] = x2.x; xv[BC - 1] = x2.y;
#pragma unroll
      for (int rr = 0; rr < BR; ++rr) { // (row start + rr*L + 2b) is even: both codes in one 32-bit load
        const unsigned pr = *reinterpret_cast<const unsigned *>(sc + rr * L + b * 2);
        av[rr][0] = sd[pr & 0xffffu]; av[rr][BC - 1] = sd[pr >> 16];
      }
    } else {
      xv[0] = xs.load(c0);
#pragma unroll
      for (int rr = 0; rr < BR; ++rr) av[rr][0] = sd[sc[rr * L + b]];
    }
  };
  int it = 0;
  for (int idx = blockIdx.x; idx < ntiles; idx += gridDim.x, ++it) {
    const int tile = tile_list ? tile_list[idx] : idx;
    if (!waited && idx >= n_nowait) { halo_wait_cta(xs); waited = true; }
    const int stage = it % stages;
    const unsigned parity = (unsigned)(it / stages) & 1u;
    const int I = tile * NB + tid;
    const int g_al = bptr[tile * NB];
    const int bs = bptr[I < nbrows ? I : nbrows];
    const int be = bptr[I + 1 < nbrows ? I + 1 : nbrows];
    mbar_wait(&full[stage], parity);
    const unsigned char *base = s_raw + stage_bytes * stage;
    const double *sd = reinterpret_cast<const double *>(base);
    const unsigned short *sc = reinterpret_cast<const unsigned short *>(base + off_codes) + (bs * (BR * BC) - ((g_al * (BR * BC)) & ~7));
    const int *sb = reinterpret_cast<const int *>(base + off_bcol) + (bs - (g_al & ~3));
    if (I < nbrows) {
      double sum[BR];
#pragma unroll
      for (int rr = 0; rr < BR; ++rr) sum[rr] = 0.0;
      const int nb = be - bs, L = nb * BC;
      int b = 0;
      for (; b + UB <= nb; b += UB) {
        double av[UB][BR][BC], xv[UB][BC];
#pragma unroll
        for (int u = 0; u < UB; ++u) fetch(sd, sc, sb, L, b + u, av[u], xv[u]);
#pragma unroll
        for (int u = 0; u < UB; ++u)
#pragma unroll
          for (int rr = 0; rr < BR; ++rr)
#pragma unroll
            for (int cc = 0; cc < BC; ++cc) sum[rr] += av[u][rr][cc] * xv[u][cc];
      }
      if (b < nb) { // tail under predicates
        double av[UB - 1][BR][BC], xv[UB - 1][BC];
#pragma unroll
        for (int u = 0; u < UB - 1; ++u)
          if (b + u < nb) fetch(sd, sc, sb, L, b + u, av[u], xv[u]);
#pragma unroll
        for (int u = 0; u < UB - 1; ++u)
          if (b + u < nb) {
#pragma unroll
            for (int rr = 0; rr < BR; ++rr)
#pragma unroll
              for (int cc = 0; cc < BC; ++cc) sum[rr] += av[u][rr][cc] * xv[u][cc];
          }
      }
      if (BR == 2 && epi.vec2) {
        double v0, v1;
        epi.apply2_store(sum[0], sum[BR - 1], I * BR, y, v0, v1);
        if (epi.push.grp) {
          if (epi.push.dof == 2) epi.push.node2(I, v0, v1);
          else { epi.push.row(I * BR, v0); epi.push.row(I * BR + 1, v1); }
        }
      } else {
#pragma unroll
        for (int rr = 0; rr < BR; ++rr) {
          const double v = epi.apply(sum[rr], I * BR + rr);
          y[I * BR + rr] = v;
          if (epi.push.grp) epi.push.row(I * BR + rr, v);
        }
      }
    }
    __syncthreads();
    if (tid == 0) {
      const int t = idx + stages * gridDim.x;
      if (t < ntiles) issue(t, stage);
    }
  }
  if (epi.push.grp) epi.push.finish(); // fused halo push: fence, ticket, last CTA raises the neighbours' flags
}

// one CTA per tile: distinct bit patterns through a shared-memory hash set; codes are ranks in order of FIRST
// OCCURRENCE (atomicMin of the position per key + a block scan), so the format is deterministic.  Run twice: the
// count pass sizes the dictionaries, the write pass fills dictionaries and codes.
__global__ void __launch_bounds__(256) k_dict_build(int nrows, int R, const int *__restrict__ rowptr, const double *__restrict__ val, int write,
                                                    int *dcnt, const int *__restrict__ dptr, double *dict, unsigned short *codes, int *stat) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  unsigned long long *keys = reinterpret_cast<unsigned long long *>(s_raw);
  int *minpos = reinterpret_cast<int *>(keys + DICT_T);
  unsigned short *rank = reinterpret_cast<unsigned short *>(minpos + DICT_T);
  unsigned short *slot = rank + DICT_T;
  __shared__ int s_scan[256];
  __shared__ int s_total;
  const int tid = threadIdx.x, tile = blockIdx.x;
  const int r0 = tile * R, r1 = min(r0 + R, nrows);
  const int j0 = rowptr[r0], n = rowptr[r1] - j0;
  if (n > DICT_MAX_TILE_NNZ) {
    if (tid == 0) { stat[0] = 1; if (!write) dcnt[tile] = 0; }
    return;
  }
  for (int h = tid; h < DICT_T; h += 256) { keys[h] = DICT_EMPTY; minpos[h] = 0x7fffffff; }
  __syncthreads();
  for (int j = tid; j < n; j += 256) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(val[j0 + j]);
    if (bits == DICT_EMPTY) { stat[0] = 1; slot[j] = 0; continue; } // the one pattern the set cannot hold: no dictionary
    unsigned h = (unsigned)((bits * 0x9E3779B97F4A7C15ull) >> 51) & (DICT_T - 1);
    for (;;) {
      const unsigned long long prev = atomicCAS(&keys[h], DICT_EMPTY, bits);
      if (prev == DICT_EMPTY || prev == bits) break;
      h = (h + 1) & (DICT_T - 1);
    }
    atomicMin(&minpos[h], j);
    slot[j] = (unsigned short)h;
  }
  __syncthreads();
  const int C = (n + 255) / 256;
  const int b = min(tid * C, n), e = min(b + C, n);
  int cnt = 0;
  for (int j = b; j < e; ++j) cnt += minpos[slot[j]] == j;
  s_scan[tid] = cnt;
  __syncthreads();
  if (tid == 0) {
    int acc = 0;
    for (int t = 0; t < 256; ++t) { const int v = s_scan[t]; s_scan[t] = acc; acc += v; }
    s_total = acc;
  }
  __syncthreads();
  const int total = s_total;
  if (!write) {
    if (tid == 0) {
      dcnt[tile] = (total + 1) & ~1; // 16-byte granules for the bulk copy
      atomicMax(&stat[1], total);
      if (total > DICT_MAX_ENTRIES) stat[0] = 1;
    }
    return;
  }
  int at = s_scan[tid];
  const int d0 = dptr[tile];
  for (int j = b; j < e; ++j)
    if (minpos[slot[j]] == j) { rank[slot[j]] = (unsigned short)at; dict[d0 + at] = val[j0 + j]; ++at; }
  if (tid == 0 && (total & 1)) dict[d0 + total] = 0.0;
  __syncthreads();
  for (int j = tid; j < n; j += 256) codes[j0 + j] = rank[slot[j]];
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

// tiles of T rows that read no ghost column come first: a kernel that waits for the neighbours' halo itself does so
// only when it reaches the first tile that needs it, so the wait hides behind the interior tiles
__global__ void __launch_bounds__(256) k_tile_ghost_flag(int nrows, int T, int ncols, const int *__restrict__ rowptr, const int *__restrict__ col, int *flag) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x) {
    bool g = false;
    for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) g = g || col[k] >= ncols;
    if (g) flag[r / T] = 1;
  }
}
static void ensure_wait_order(const Csr &A, int T) {
  static const bool off = getenv("B200SP_NO_DEFER_WAIT") && atoi(getenv("B200SP_NO_DEFER_WAIT"));
  if (off || A.wait_order_rows == T || A.nrows <= 0) return;
  Ctx *c = A.ctx;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  B2_CUDA(cudaStreamIsCapturing(c->stream, &st));
  if (st != cudaStreamCaptureStatusNone) return;
  const int ntiles = (A.nrows + T - 1) / T;
  DevBuf<int> flag((size_t)ntiles + 1);
  flag.zero(c->stream);
  {
    LaunchScope ls(c, "setup");
    k_tile_ghost_flag<<<std::max(1, std::min((A.nrows + 255) / 256, c->num_sms * 16)), 256, 0, c->stream>>>(A.nrows, T, A.ncols, A.rowptr.p, A.col.p, flag.p);
    check_launch("k_tile_ghost_flag");
  }
  std::vector<int> hf((size_t)ntiles), order;
  B2_CUDA(cudaMemcpyAsync(hf.data(), flag.p, sizeof(int) * (size_t)ntiles, cudaMemcpyDeviceToHost, c->stream));
  c->sync();
  order.reserve((size_t)ntiles);
  for (int t = 0; t < ntiles; ++t) if (!hf[(size_t)t]) order.push_back(t);
  A.wait_n_nowait = (int)order.size();
  for (int t = 0; t < ntiles; ++t) if (hf[(size_t)t]) order.push_back(t);
  A.wait_order.alloc((size_t)ntiles + 1);
  B2_CUDA(cudaMemcpyAsync(A.wait_order.p, order.data(), sizeof(int) * (size_t)ntiles, cudaMemcpyHostToDevice, c->stream));
  c->sync();
  A.wait_order_rows = T;
}

// build (or decline) the tile-local value dictionary of A; called lazily from the first un-captured TMA SpMV
static void build_value_dict(const Csr &A) {
  static const bool off = getenv("B200SP_NO_VALUE_DICT") && atoi(getenv("B200SP_NO_VALUE_DICT"));
  Ctx *c = A.ctx;
  A.dict_state = -1;
  if (off || A.nrows == 0 || A.nnz < 4096 || spmv_tma_tile_rows() != TMA_TILE_ROWS) return;
  const int R = TMA_TILE_ROWS * (A.bcol.p ? A.blk_r : 1), ntiles = (A.nrows + R - 1) / R; // one thread per block row
  A.dict_rows = R;
  DevBuf<int> cnt((size_t)ntiles + 1), stat(2);
  stat.zero(c->stream);
  const size_t smem = (size_t)DICT_T * 14 + (size_t)DICT_MAX_TILE_NNZ * 2;
  static bool attr = false;
  if (!attr) { B2_CUDA(cudaFuncSetAttribute(k_dict_build, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
  { LaunchScope ls(c, "setup"); k_dict_build<<<ntiles, 256, smem, c->stream>>>(A.nrows, R, A.rowptr.p, A.val.p, 0, cnt.p, nullptr, nullptr, nullptr, stat.p); check_launch("k_dict_build"); }
  int h_stat[2] = {0, 0};
  B2_CUDA(cudaMemcpyAsync(h_stat, stat.p, sizeof(h_stat), cudaMemcpyDeviceToHost, c->stream));
  c->sync();
  if (h_stat[0]) return;
  A.dptr.alloc((size_t)ntiles + 1);
  int total = 0;
  exclusive_scan_i32(c, cnt.p, A.dptr.p, ntiles, &total);
  // worth it only if dictionary + codes are clearly smaller than the 8 B/nnz value stream
  if ((double)total * 8.0 + (double)A.nnz * 2.0 > 0.75 * 8.0 * (double)A.nnz) { A.dptr.release(); return; }
  A.dict.alloc((size_t)total + 16);
  A.codes.alloc((size_t)A.nnz + 32);
  A.codes.zero(c->stream);
  { LaunchScope ls(c, "setup"); k_dict_build<<<ntiles, 256, smem, c->stream>>>(A.nrows, R, A.rowptr.p, A.val.p, 1, nullptr, A.dptr.p, A.dict.p, A.codes.p, stat.p); check_launch("k_dict_build"); }
  c->sync();
  // Measured (profiles/r01_value_dict_sweep.txt): the dictionary wins whenever the block-row kernel applies (A 2x2,
  // B^T 2x1, B 1x2: 1.7-1.9x over the plain value stream), for long scalar rows, and for tiny dictionaries (R: the
  // lookup is a shared-memory broadcast); for scalar matrices with 9 nnz/row and a few hundred distinct values per
  // tile (C, Q) the extra dependent shared-memory lookup costs more than the smaller stream saves.
  static const int force = getenv("B200SP_VALUE_DICT") ? atoi(getenv("B200SP_VALUE_DICT")) : 0;
  if (force < 2 && !A.bcol.p && (double)A.nnz / A.nrows < 12.0 && h_stat[1] > 32) { A.dptr.release(); A.dict.release(); A.codes.release(); return; }
  A.dict_cap = (h_stat[1] + 1) & ~1;
  if (A.dict_cap < 2) A.dict_cap = 2;
  A.dict_bytes = (int64_t)total * 8;
  A.dict_state = 1;
}
void csr_drop_value_dict(Csr &A) {
  A.dict.release(); A.dptr.release(); A.codes.release();
  A.dict_state = 0; A.dict_cap = 0; A.dict_rows = 0; A.dict_bytes = 0;
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
    B2_CUDA(cudaFuncSetAttribute(k_spmv_tma<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    B2_CUDA(cudaFuncSetAttribute(k_spmv_tma<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    attr_set = true;
  }
  if (tile_list && R != TMA_TILE_ROWS) return false; // the lists were built for TMA_TILE_ROWS-row tiles
  const int ntiles = tile_list ? nlist : (A.nrows + R - 1) / R;
  if (ntiles <= 0) return true;
  int per_sm = (int)(budget / smem);
  if (per_sm < 1) per_sm = 1;
  if (per_sm * R > 2048) per_sm = 2048 / R;
  int grid = ntiles < c->num_sms * per_sm ? ntiles : c->num_sms * per_sm;
  if (A.dict_state == 0) { // lazily, never while a CUDA graph is being recorded
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    B2_CUDA(cudaStreamIsCapturing(c->stream, &st));
    if (st == cudaStreamCaptureStatusNone) build_value_dict(A);
  }
  const int dbr = A.bcol.p ? A.blk_r : 1, dbc = A.bcol.p ? A.blk_c : 1;
  if (A.dict_state == 1 && R == TMA_TILE_ROWS && !tile_list && A.dict_rows == TMA_TILE_ROWS * dbr &&
      (dbc == 1 || (reinterpret_cast<uintptr_t>(xs.x) & 15) == 0)) { // tile-local value dictionary, one thread per block row
    const int capt = cap * dbr;                       // nonzeros per tile of 128 block rows
    const int capc = (capt + 16) & ~7;
    const int capb = ((capt / (dbr * dbc) + 8) + 3) & ~3;
    const size_t smem_d = ((size_t)A.dict_cap * 8 + (size_t)capc * 2 + (size_t)capb * 4) * stages + 8 * TMA_MAX_STAGES;
    if (smem_d <= budget) {
      static bool attr_d = false;
      if (!attr_d) {
        B2_CUDA(cudaFuncSetAttribute(k_spmv_tma_dict<2, 2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
        B2_CUDA(cudaFuncSetAttribute(k_spmv_tma_dict<2, 1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
        B2_CUDA(cudaFuncSetAttribute(k_spmv_tma_dict<1, 2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
        B2_CUDA(cudaFuncSetAttribute(k_spmv_tma_dict<1, 1, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
        attr_d = true;
      }
      int psm = (int)(budget / smem_d);
      if (psm < 1) psm = 1;
      if (psm * R > 2048) psm = 2048 / R;
      static const int env_psm = getenv("B200SP_TMA_CTAS") ? atoi(getenv("B200SP_TMA_CTAS")) : 0;
      if (env_psm && env_psm < psm) psm = env_psm;
      const int nbrows = A.nrows / dbr;
      const int ntd = (nbrows + R - 1) / R;
      const int gridd = ntd < c->num_sms * psm ? ntd : c->num_sms * psm;
      const int *bp = A.bcol.p ? A.bptr.p : A.rowptr.p, *bc = A.bcol.p ? A.bcol.p : A.col.p;
      const int *order = nullptr;
      int n_nowait = 0;
      if (xs.wait_flags) { ensure_wait_order(A, R * dbr); if (A.wait_order_rows == R * dbr) { order = A.wait_order.p; n_nowait = A.wait_n_nowait; } }
      auto al16 = [](const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
      SpmvEpi e2 = epi;
      e2.vec2 = dbr == 2 && al16(y) && al16(epi.z) && al16(epi.pm1) && al16(epi.pk) && al16(epi.dinv);
#define B200SP_DICT_LAUNCH(BR_, BC_, UB_) \
  k_spmv_tma_dict<BR_, BC_, UB_><<<gridd, R, smem_d, c->stream>>>(nbrows, ntd, order, bp, bc, A.dptr.p, A.dict.p, A.codes.p, xs, y, e2, capc, capb, A.dict_cap, stages, n_nowait)
      if (dbr == 2 && dbc == 2) B200SP_DICT_LAUNCH(2, 2, 3);
      else if (dbr == 2 && dbc == 1) B200SP_DICT_LAUNCH(2, 1, 3);
      else if (dbr == 1 && dbc == 2) B200SP_DICT_LAUNCH(1, 2, 3);
      else B200SP_DICT_LAUNCH(1, 1, 6);
#undef B200SP_DICT_LAUNCH
      check_launch("k_spmv_tma_dict");
      return true;
    }
  }
  int n_nowait = 0;
  if (xs.wait_flags && !tile_list) { ensure_wait_order(A, R); if (A.wait_order_rows == R) { tile_list = A.wait_order.p; n_nowait = A.wait_n_nowait; } }
  if (A.bcol.p) { // block-compressed column index: 8 + 4/(BR*BC) bytes per nonzero (R is a multiple of BR)
    const int capb = ((cap / (A.blk_r * A.blk_c) + 8) + 3) & ~3;
    const size_t smem_b = ((size_t)cap * 8 + (size_t)capb * 4) * stages + 8 * TMA_MAX_STAGES;
    static bool attr_b = false;
    if (!attr_b) {
      B2_CUDA(cudaFuncSetAttribute(k_spmv_tma_blk<2, 2, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
      B2_CUDA(cudaFuncSetAttribute(k_spmv_tma_blk<1, 2, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
      B2_CUDA(cudaFuncSetAttribute(k_spmv_tma_blk<2, 1, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
      attr_b = true;
    }
    int psm = (int)(budget / smem_b);
    if (psm < 1) psm = 1;
    if (psm * R > 2048) psm = 2048 / R;
    const int gridb = ntiles < c->num_sms * psm ? ntiles : c->num_sms * psm;
    if (A.blk_r == 2 && A.blk_c == 2)
      k_spmv_tma_blk<2, 2, 6><<<gridb, R, smem_b, c->stream>>>(A.nrows, ntiles, tile_list, A.rowptr.p, A.bptr.p, A.bcol.p, A.val.p, xs, y, epi, cap, capb, stages, n_nowait);
    else if (A.blk_r == 1 && A.blk_c == 2)
      k_spmv_tma_blk<1, 2, 6><<<gridb, R, smem_b, c->stream>>>(A.nrows, ntiles, tile_list, A.rowptr.p, A.bptr.p, A.bcol.p, A.val.p, xs, y, epi, cap, capb, stages, n_nowait);
    else
      k_spmv_tma_blk<2, 1, 6><<<gridb, R, smem_b, c->stream>>>(A.nrows, ntiles, tile_list, A.rowptr.p, A.bptr.p, A.bcol.p, A.val.p, xs, y, epi, cap, capb, stages, n_nowait);
    check_launch("k_spmv_tma_blk");
    return true;
  }
  if (env_U ? env_U == 6 : true)
    k_spmv_tma<6><<<grid, R, smem, c->stream>>>(A.nrows, ntiles, tile_list, A.rowptr.p, A.col.p, A.val.p, xs, y, epi, cap, stages, n_nowait);
  else
    k_spmv_tma<3><<<grid, R, smem, c->stream>>>(A.nrows, ntiles, tile_list, A.rowptr.p, A.col.p, A.val.p, xs, y, epi, cap, stages, n_nowait);
  check_launch("k_spmv_tma");
  return true;
}

} // namespace b200sp
