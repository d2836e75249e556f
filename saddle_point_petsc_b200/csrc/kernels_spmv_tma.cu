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

// ---- tile-local pattern / value dictionaries ("pd" format) ---------------------------------------------------------
// Finite-element matrices on structured grids repeat themselves: inside a tile of 128 block rows (nodes) the COLUMN
// PATTERN of a row (its block columns relative to the row) takes one or a few values, and at a given position of the
// stencil the VALUE takes a handful of distinct bit patterns (they are not all equal because (i+1)h - ih rounds
// differently from element to element).  A tile is therefore stored as ONE contiguous blob
//     header | dict: distinct values, grouped BY POSITION (block k, entry (rr,cc)), order of first occurrence
//            | patterns: {number of blocks, block-column offsets}  | posoff: start of every position's group (u16)
//            | pid: one pattern id per block row (u8)               | codes: [block k][row][BR*BC] one byte per nonzero
// and the kernel streams ~1 byte per nonzero instead of 12 (A block: 0.58 GB instead of 2.5 GB per MatMult; the
// previous tile-global 16-bit dictionary + explicit block columns moved 0.88 GB).  The value fed to the multiply is the
// identical double, in the identical CSR order, so the result is bit-for-bit that of every other SpMV kernel.
// Two things make the lookups cheap: codes and pattern ids are stored ELL-style inside the tile (consecutive lanes read
// consecutive bytes: no bank conflicts), and because the dictionary is grouped by position, the 32 lanes of a warp --
// which look up the SAME position for neighbouring nodes -- hit a few CONSECUTIVE entries (broadcast or distinct banks)
// instead of random entries of a tile-wide dictionary (the measured limiter of the previous format: 40 % of the
// shared-memory wavefronts were bank-conflict replays).  Matrices whose tiles do not compress keep the plain streams.
constexpr int PD_NB = 128;         // block rows per tile (= blockDim.x of the SpMV kernel)
constexpr int PD_MAX_K = 16;       // blocks per block row (node-block matrices)
constexpr int PD_MAX_K_SCALAR = 96; // entries per row of a scalar (1 x 1) matrix: covers the 81 of the 3-D operators
template <int BR, int BC> __host__ __device__ constexpr int pd_max_k() { return BR * BC == 1 ? PD_MAX_K_SCALAR : PD_MAX_K; }
constexpr int PD_MAX_DICT = 3072;  // (position, value) pairs per tile

__host__ __device__ inline int pd_pad16(int bytes) { return (bytes + 15) & ~15; }
// rel: what the pattern's block-column offsets are relative to.  0: the block row index itself (square stencil matrices:
// one pattern per tile); 1: the row's FIRST block column, stored explicitly per row (+4 bytes per row) -- for rectangular
// operators whose columns move at another rate than the rows (interpolation / restriction: 4 resp. 1 patterns per tile
// instead of 128).
struct PdLayout { int nbmax, npat, ndict, rel, o_dict, o_pat, o_pos, o_pid, o_base, o_codes, bytes; };
template <int BR, int BC>
__host__ __device__ inline PdLayout pd_layout(int nbmax, int npat, int ndict, int rel) {
  PdLayout L;
  L.nbmax = nbmax; L.npat = npat; L.ndict = ndict; L.rel = rel;
  L.o_dict = 16;
  L.o_pat = L.o_dict + pd_pad16(ndict * 8);
  L.o_pos = L.o_pat + pd_pad16(npat * (nbmax + 1) * 4);
  L.o_pid = L.o_pos + pd_pad16(nbmax * BR * BC * 2);
  L.o_base = L.o_pid + PD_NB;
  L.o_codes = L.o_base + (rel ? PD_NB * 4 : 0);
  L.bytes = L.o_codes + nbmax * PD_NB * BR * BC;
  return L;
}

// One thread per BLOCK ROW (node): one x load (16 bytes when BC = 2) serves BR x BC nonzeros; each row accumulates its
// own entries in CSR order (block k ascending, then column).  One bulk copy per tile, two stages per CTA.
template <int BR, int BC, int UB>
__global__ void __launch_bounds__(PD_NB) k_spmv_pd(int nbrows, const int *__restrict__ gstart, int ntiles, const int *__restrict__ order, const int *__restrict__ toff,
                                                   const unsigned char *__restrict__ blob, XSrc xs, double *y, SpmvEpi epi, int cap, int stages,
                                                   int n_nowait) {
  // A stage holds a GROUP of G consecutive tiles (their blobs are contiguous in memory: one bulk copy): matrices with short
  // rows have blobs of 1-4 KB, and one tile per stage left the kernel bound by the turn-around latency of the copies.
  // `ntiles` counts groups; group g covers tiles [gstart[g], gstart[g+1]) -- formed on the host by bytes (~8 KB), so that one
  // unusually large tile (domain boundary, ghost columns) does not inflate the stage size of every CTA.
  extern __shared__ __align__(128) unsigned char s_raw[];
  constexpr int BRBC = BR * BC;
  uint64_t *full = reinterpret_cast<uint64_t *>(s_raw + (size_t)cap * stages);
  const int tid = threadIdx.x;
  auto issue = [&](int idx, int stage) { // one thread
    const int grp = order ? order[idx] : idx;
    const long long o0 = toff[grp], o1 = toff[grp + 1]; // toff: blob offset of every GROUP (16-byte units)
    const unsigned bytes = (unsigned)(o1 - o0) * 16u;
    mbar_expect_tx(&full[stage], bytes);
    if (bytes) bulk_g2s(s_raw + (size_t)cap * stage, blob + o0 * 16, bytes, &full[stage]);
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
    const int grp = order ? order[idx] : idx;
    if (!waited && idx >= n_nowait) { halo_wait_cta(xs); waited = true; }
    const int stage = it % stages;
    const unsigned parity = (unsigned)(it / stages) & 1u;
    mbar_wait(&full[stage], parity);
    const unsigned char *base = s_raw + (size_t)cap * stage;
    const int t_end = gstart[grp + 1];
    for (int tile = gstart[grp]; tile < t_end; ++tile) {
     const int I = tile * PD_NB + tid;
     const int4 hdr = *reinterpret_cast<const int4 *>(base);
     if (I < nbrows) {
      const PdLayout L = pd_layout<BR, BC>(hdr.x, hdr.y, hdr.z, hdr.w);
      const int cbase = hdr.w ? reinterpret_cast<const int *>(base + L.o_base)[tid] : I; // what the pattern offsets are relative to
      const double *sd = reinterpret_cast<const double *>(base + L.o_dict);
      const int *sp = reinterpret_cast<const int *>(base + L.o_pat) + (int)base[L.o_pid + tid] * (L.nbmax + 1);
      const unsigned short *spos = reinterpret_cast<const unsigned short *>(base + L.o_pos);
      const unsigned char *sc = base + L.o_codes + tid * BRBC;
      const int nb = sp[0];
      // values of block k of this thread's rows and the x entries they multiply
      auto fetch = [&](int k, double (&av)[BR][BC], double (&xv)[BC]) {
        const int c0 = (cbase + sp[1 + k]) * BC;
        if (BC == 2) {
          const double2 x2 = xs.load2(c0);
          xv[0] = x2.x; xv[BC - 1] = x2.y;
        } else {
          xv[0] = xs.load(c0);
        }
        if (BRBC == 4) {
          const unsigned cw = *reinterpret_cast<const unsigned *>(sc + (size_t)k * (PD_NB * 4));
          const uint2 po = *reinterpret_cast<const uint2 *>(spos + k * 4);
          av[0][0] = sd[(po.x & 0xffffu) + (cw & 0xffu)];
          av[0][BC - 1] = sd[(po.x >> 16) + ((cw >> 8) & 0xffu)];
          av[BR - 1][0] = sd[(po.y & 0xffffu) + ((cw >> 16) & 0xffu)];
          av[BR - 1][BC - 1] = sd[(po.y >> 16) + (cw >> 24)];
        } else if (BRBC == 2) {
          const unsigned cw = *reinterpret_cast<const unsigned short *>(sc + (size_t)k * (PD_NB * 2));
          const unsigned po = *reinterpret_cast<const unsigned *>(spos + k * 2);
          av[0][0] = sd[(po & 0xffffu) + (cw & 0xffu)];
          av[BR - 1][BC - 1] = sd[(po >> 16) + (cw >> 8)];
        } else {
          av[0][0] = sd[(unsigned)spos[k] + (unsigned)sc[(size_t)k * PD_NB]];
        }
      };
      double sum[BR];
#pragma unroll
      for (int rr = 0; rr < BR; ++rr) sum[rr] = 0.0;
      int b = 0;
      for (; b + UB <= nb; b += UB) {
        double av[UB][BR][BC], xv[UB][BC];
#pragma unroll
        for (int u = 0; u < UB; ++u) fetch(b + u, av[u], xv[u]);
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
          if (b + u < nb) fetch(b + u, av[u], xv[u]);
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
     base += pd_layout<BR, BC>(hdr.x, hdr.y, hdr.z, hdr.w).bytes; // the next tile of the group follows immediately
    }
    __syncthreads();
    if (tid == 0) {
      const int t = idx + stages * gridDim.x;
      if (t < ntiles) issue(t, stage);
    }
  }
  if (epi.push.grp) epi.push.finish(); // fused halo push: fence, ticket, last CTA raises the neighbours' flags
}

// One CTA per tile builds (write = 0: sizes only) the blob described above.  Everything is decided by position in the
// tile (first occurrence wins, ranks are prefix counts), so the format is deterministic.  Reads the CSR arrays; the
// block structure (BR rows of a node share their columns, columns come in aligned runs of BC) was verified by
// k_blk_check when BR*BC > 1.
template <int BR, int BC>
__global__ void __launch_bounds__(256) k_pd_build(int nbrows, const int *__restrict__ rowptr, const int *__restrict__ col, const double *__restrict__ val,
                                                  int write, int rel, int KS, int *tsize16, const int *__restrict__ toff, unsigned char *blob, int *stat) {
  // KS: capacity in blocks per row of the shared arrays (>= the longest row of THIS matrix, odd so that KS + 1 is even and
  // the 8-byte array behind s_delta stays aligned); the static per-position arrays are sized for the template's maximum
  extern __shared__ __align__(16) unsigned char s_raw[];
  constexpr int BRBC = BR * BC, PMAX = pd_max_k<BR, BC>() * BRBC;
  const int PS = KS * BRBC;
  int *s_delta = reinterpret_cast<int *>(s_raw);                                                   // [PD_NB][KS + 1]: nb, offsets
  unsigned long long *s_val = reinterpret_cast<unsigned long long *>(s_delta + PD_NB * (KS + 1));  // [PS][PD_NB] value bits
  unsigned char *s_first = reinterpret_cast<unsigned char *>(s_val + (size_t)PS * PD_NB);           // [PS][PD_NB] first row with this value
  unsigned char *s_rank = s_first + (size_t)PS * PD_NB;                                            // [PS][PD_NB] rank of a first occurrence
  __shared__ int s_pfirst[PD_NB], s_prank[PD_NB], s_cnt[PMAX], s_posoff[PMAX + 1], s_w[4], s_base[PD_NB];
  __shared__ unsigned long long s_hkey[8][256];
  __shared__ int s_hmin[8][256];
  __shared__ int s_nbmax, s_bad, s_npat;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, tile = blockIdx.x;
  const int I0 = tile * PD_NB;
  if (tid == 0) { s_nbmax = 0; s_bad = 0; }
  __syncthreads();
  if (tid < PD_NB) {
    const int I = I0 + tid;
    int nb = 0;
    if (I < nbrows) {
      const int rs = rowptr[I * BR];
      nb = (rowptr[I * BR + 1] - rs) / BC;
      if (nb > KS) { s_bad = 1; nb = 0; }
      const int ref = rel ? (nb > 0 ? col[rs] / BC : 0) : I;
      s_base[tid] = ref;
      for (int k = 0; k < nb; ++k) s_delta[tid * (KS + 1) + 1 + k] = col[rs + k * BC] / BC - ref;
    } else {
      s_base[tid] = 0;
    }
    s_delta[tid * (KS + 1)] = nb;
    atomicMax(&s_nbmax, nb);
  }
  __syncthreads();
  if (s_bad) {
    if (tid == 0) { stat[0] = 1; if (!write) tsize16[tile] = 0; }
    return;
  }
  const int nbmax = s_nbmax, P = nbmax * BRBC;
  // value bits by position
  for (int idx = tid; idx < P * PD_NB; idx += 256) {
    const int p = idx / PD_NB, i = idx % PD_NB, k = p / BRBC, e = p % BRBC, rr = e / BC, cc = e % BC;
    unsigned long long bits = 0ull;
    if (k < s_delta[i * (KS + 1)]) bits = (unsigned long long)__double_as_longlong(val[rowptr[(I0 + i) * BR + rr] + k * BC + cc]);
    s_val[idx] = bits;
  }
  // column patterns: first row with the same {nb, offsets}
  if (tid < PD_NB) {
    const int *mine = s_delta + tid * (KS + 1);
    int first = tid;
    for (int j = 0; j < tid; ++j) {
      const int *o = s_delta + j * (KS + 1);
      bool same = o[0] == mine[0];
      for (int k = 0; k < mine[0] && same; ++k) same = o[1 + k] == mine[1 + k];
      if (same) { first = j; break; }
    }
    s_pfirst[tid] = first;
  }
  __syncthreads();
  if (tid < PD_NB) { // rank of the first occurrences (prefix count over the tile)
    const bool isf = s_pfirst[tid] == tid;
    const unsigned bal = __ballot_sync(FULL, isf);
    if (lane == 0) s_w[warp] = __popc(bal);
    s_prank[tid] = __popc(bal & ((1u << lane) - 1u)); // completed below with the warps before this one
  }
  __syncthreads();
  if (tid < PD_NB) {
    int base = 0;
    for (int w = 0; w < warp; ++w) base += s_w[w];
    s_prank[tid] += base;
    if (tid == 0) s_npat = s_w[0] + s_w[1] + s_w[2] + s_w[3];
  }
  // values: first row of the tile that holds the same bits at the same position.  One warp per position, a 256-slot
  // hash set of the position's (at most 128) values in shared memory: insert with atomicCAS, keep the smallest row per
  // key with atomicMin -- the result (first occurrence) does not depend on the order of insertion.
  for (int p = warp; p < P; p += 8) {
    unsigned long long *hk = s_hkey[warp];
    int *hm = s_hmin[warp];
    for (int q = lane; q < 256; q += 32) { hk[q] = ~0ull; hm[q] = 0x7fffffff; }
    __syncwarp();
    const int k = p / BRBC;
    int slot[PD_NB / 32];
#pragma unroll
    for (int c = 0; c < PD_NB / 32; ++c) {
      const int i = c * 32 + lane;
      slot[c] = -1;
      if (k < s_delta[i * (KS + 1)]) {
        const unsigned long long v = s_val[p * PD_NB + i];
        if (v == ~0ull) { s_bad = 1; continue; } // the one bit pattern the set cannot hold (a NaN payload): no dictionary
        unsigned h = (unsigned)((v * 0x9E3779B97F4A7C15ull) >> 56);
        for (;;) {
          const unsigned long long prev = atomicCAS(&hk[h], ~0ull, v);
          if (prev == ~0ull || prev == v) break;
          h = (h + 1) & 255u;
        }
        atomicMin(&hm[h], i);
        slot[c] = (int)h;
      }
    }
    __syncwarp();
#pragma unroll
    for (int c = 0; c < PD_NB / 32; ++c) {
      const int i = c * 32 + lane;
      s_first[p * PD_NB + i] = (unsigned char)(slot[c] >= 0 ? hm[slot[c]] : 255);
    }
    __syncwarp();
  }
  __syncthreads();
  for (int p = warp; p < P; p += 8) { // one warp per position: ranks of the first occurrences, in row order
    int base = 0;
    for (int c = 0; c < PD_NB / 32; ++c) {
      const int i = c * 32 + lane;
      const bool isf = s_first[p * PD_NB + i] == i;
      const unsigned bal = __ballot_sync(FULL, isf);
      s_rank[p * PD_NB + i] = (unsigned char)(base + __popc(bal & ((1u << lane) - 1u)));
      base += __popc(bal);
    }
    if (lane == 0) s_cnt[p] = base;
  }
  __syncthreads();
  if (tid == 0) {
    int acc = 0;
    for (int p = 0; p < P; ++p) { s_posoff[p] = acc; acc += s_cnt[p]; }
    s_posoff[P] = acc;
    if (acc > PD_MAX_DICT) s_bad = 1;
  }
  __syncthreads();
  const int ndict = s_posoff[P], npat = s_npat;
  if (s_bad) {
    if (tid == 0) { stat[0] = 1; if (!write) tsize16[tile] = 0; }
    return;
  }
  const PdLayout L = pd_layout<BR, BC>(nbmax, npat, ndict, rel);
  if (!write) {
    if (tid == 0) { tsize16[tile] = L.bytes / 16; atomicMax(&stat[1], L.bytes); }
    return;
  }
  unsigned char *out = blob + (size_t)toff[tile] * 16; // zero-filled by the caller: padding stays zero
  if (tid == 0) { int *h = reinterpret_cast<int *>(out); h[0] = nbmax; h[1] = npat; h[2] = ndict; h[3] = rel; }
  if (rel && tid < PD_NB) reinterpret_cast<int *>(out + L.o_base)[tid] = s_base[tid];
  double *dict = reinterpret_cast<double *>(out + L.o_dict);
  int *pat = reinterpret_cast<int *>(out + L.o_pat);
  unsigned short *pos = reinterpret_cast<unsigned short *>(out + L.o_pos);
  for (int idx = tid; idx < P * PD_NB; idx += 256) {
    const int p = idx / PD_NB, i = idx % PD_NB, k = p / BRBC, e = p % BRBC;
    const int first = s_first[idx];
    if (first == i) dict[s_posoff[p] + s_rank[idx]] = __longlong_as_double((long long)s_val[idx]);
    out[L.o_codes + ((size_t)k * PD_NB + i) * BRBC + e] = first == 255 ? (unsigned char)0 : s_rank[p * PD_NB + first];
  }
  if (tid < PD_NB) {
    out[L.o_pid + tid] = (unsigned char)s_prank[s_pfirst[tid]];
    if (s_pfirst[tid] == tid) {
      const int *mine = s_delta + tid * (KS + 1);
      int *q = pat + s_prank[tid] * (nbmax + 1);
      q[0] = mine[0];
      for (int k = 0; k < nbmax; ++k) q[1 + k] = k < mine[0] ? mine[1 + k] : 0;
    }
  }
  for (int p = tid; p < P; p += 256) pos[p] = (unsigned short)s_posoff[p];
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
  if (off) return false;
  if (br * bc >= 3 && (br == 3 || bc == 3) && br * bc != 6) { // 3-D node blocks: warp-per-node kernel (kernels_spmv.cu k_spmv_nodeblk)
    if (A.kernel == SPMV_TMA) return false; // short rows (B^T: 27 per row): the TMA thread-per-row kernel is faster (0.62 vs 0.70 ms at 8.6M DOF)
    bool ok = false;
    if (br == 3 && bc == 3) ok = build_block_index<3, 3>(A);
    else if (br == 3 && bc == 1) ok = build_block_index<3, 1>(A);
    else if (br == 1 && bc == 3) ok = build_block_index<1, 3>(A);
    if (ok) A.kernel = SPMV_NODE;
    return ok;
  }
  if (A.kernel != SPMV_TMA) return false;
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

// the same ordering for the pd kernel's GROUPS of tiles (variable size): a group waits if any of its tiles has a ghost column.
// Cached as wait_order_rows = -T so that it is not confused with the fixed-size tile order of the other kernels.
static void ensure_wait_order_groups(const Csr &A, int T) {
  static const bool off = getenv("B200SP_NO_DEFER_WAIT") && atoi(getenv("B200SP_NO_DEFER_WAIT"));
  if (off || A.wait_order_rows == -T || A.nrows <= 0 || A.h_pd_gstart.empty()) return;
  Ctx *c = A.ctx;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  B2_CUDA(cudaStreamIsCapturing(c->stream, &st));
  if (st != cudaStreamCaptureStatusNone) return;
  const int ntiles = (A.nrows + T - 1) / T, ng = (int)A.h_pd_gstart.size() - 1;
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
  std::vector<char> gf((size_t)ng, 0);
  for (int g = 0; g < ng; ++g)
    for (int t = A.h_pd_gstart[(size_t)g]; t < A.h_pd_gstart[(size_t)g + 1] && t < ntiles; ++t) gf[(size_t)g] |= (char)(hf[(size_t)t] != 0);
  order.reserve((size_t)ng);
  for (int g = 0; g < ng; ++g) if (!gf[(size_t)g]) order.push_back(g);
  A.wait_n_nowait = (int)order.size();
  for (int g = 0; g < ng; ++g) if (gf[(size_t)g]) order.push_back(g);
  A.wait_order.alloc((size_t)ng + 1);
  B2_CUDA(cudaMemcpyAsync(A.wait_order.p, order.data(), sizeof(int) * (size_t)ng, cudaMemcpyHostToDevice, c->stream));
  c->sync();
  A.wait_order_rows = -T;
}

// build (or decline) the tile-local pattern/value dictionaries of A; called lazily from the first un-captured TMA SpMV
template <int BR, int BC>
static bool build_pd_mode(const Csr &A, int force, int rel, bool only_if_small) {
  Ctx *c = A.ctx;
  const int nbrows = A.nrows / BR, ntiles = (nbrows + PD_NB - 1) / PD_NB;
  DevBuf<int> tsize((size_t)ntiles + 1), stat(2);
  stat.zero(c->stream);
  // capacity of the build kernel's shared arrays: the longest row of this matrix (odd), at most the template's maximum
  int KS = (A.max_row_nnz + BC - 1) / BC;
  if (KS > pd_max_k<BR, BC>()) return false;
  KS |= 1;
  const size_t smem = (size_t)PD_NB * (KS + 1) * 4 + (size_t)KS * BR * BC * PD_NB * 10;
  // (the attribute is per device and this kernel has four instantiations: set it on every build, setup path only)
  B2_CUDA(cudaFuncSetAttribute(k_pd_build<BR, BC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  { LaunchScope ls(c, "setup"); k_pd_build<BR, BC><<<ntiles, 256, smem, c->stream>>>(nbrows, A.rowptr.p, A.col.p, A.val.p, 0, rel, KS, tsize.p, nullptr, nullptr, stat.p); check_launch("k_pd_build"); }
  int h_stat[2] = {0, 0};
  B2_CUDA(cudaMemcpyAsync(h_stat, stat.p, sizeof(h_stat), cudaMemcpyDeviceToHost, c->stream));
  c->sync();
  if (h_stat[0]) return false; // a row with more than PD_MAX_K blocks, or a tile with too many distinct values
  DevBuf<int> off((size_t)ntiles + 1);
  int total16 = 0;
  exclusive_scan_i32(c, tsize.p, off.p, ntiles, &total16);
  const double total = 16.0 * (double)total16;
  // worth it only if the blobs are clearly smaller than what the plain kernels stream (values + (block) column index)
  const double alt = 8.0 * (double)A.nnz + 4.0 * (double)A.nnz / (double)(BR * BC);
  if (force < 2 && total > 0.6 * alt) return false;
  if (only_if_small && total > 0.25 * alt) return false; // the caller has a second mode to try: accept only a clear win
  A.pd_off = std::move(off);
  A.pd_blob.alloc((size_t)total16 * 16 + 256);
  A.pd_blob.zero(c->stream);
  { LaunchScope ls(c, "setup"); k_pd_build<BR, BC><<<ntiles, 256, smem, c->stream>>>(nbrows, A.rowptr.p, A.col.p, A.val.p, 1, rel, KS, nullptr, A.pd_off.p, A.pd_blob.p, stat.p); check_launch("k_pd_build"); }
  c->sync();
  { // groups of consecutive tiles, ~8 KB of blobs and at most 8 tiles each (measured: profiles/r02_pd_group_sweep.txt);
    // the stage capacity is the largest group
    std::vector<int> h_off((size_t)ntiles + 1), gs;
    B2_CUDA(cudaMemcpyAsync(h_off.data(), A.pd_off.p, sizeof(int) * ((size_t)ntiles + 1), cudaMemcpyDeviceToHost, c->stream));
    c->sync();
    static const int env_g = getenv("B200SP_PD_GROUP") ? atoi(getenv("B200SP_PD_GROUP")) : 0;       // fixed tiles per group
    static const int env_kb = getenv("B200SP_PD_GROUP_KB") ? atoi(getenv("B200SP_PD_GROUP_KB")) : 0; // byte target
    const int64_t target = env_kb > 0 ? (int64_t)env_kb * 1024 : 12288;
    int64_t capg = 0;
    for (int t = 0; t < ntiles;) {
      gs.push_back(t);
      int e = t + 1;
      if (env_g > 0) e = std::min(t + env_g, ntiles);
      else while (e < ntiles && e - t < 8 && 16LL * (h_off[(size_t)e + 1] - h_off[(size_t)t]) <= target) ++e;
      capg = std::max<int64_t>(capg, 16LL * (h_off[(size_t)e] - h_off[(size_t)t]));
      t = e;
    }
    gs.push_back(ntiles);
    A.pd_ngroups = (int)gs.size() - 1;
    A.pd_gstart.alloc(gs.size() + 1);
    A.pd_goff.alloc(gs.size() + 1);
    std::vector<int> go(gs.size());
    for (size_t g = 0; g < gs.size(); ++g) go[g] = h_off[(size_t)gs[g]];
    B2_CUDA(cudaMemcpyAsync(A.pd_gstart.p, gs.data(), sizeof(int) * gs.size(), cudaMemcpyHostToDevice, c->stream));
    B2_CUDA(cudaMemcpyAsync(A.pd_goff.p, go.data(), sizeof(int) * go.size(), cudaMemcpyHostToDevice, c->stream));
    c->sync();
    A.h_pd_gstart = gs;
    A.pd_cap = (int)((capg + 127) & ~127LL);
    static const bool dbg = getenv("B200SP_PD_DEBUG") && atoi(getenv("B200SP_PD_DEBUG"));
    if (dbg) fprintf(stderr, "[b200sp pd] %s: %dx%d blocks, rel %d, %d tiles, %.2f B/nnz, largest tile %d B, %d groups, stage %d B\n", A.tag.c_str(), BR, BC, rel,
                     ntiles, total / (double)std::max<int64_t>(A.nnz, 1), h_stat[1], A.pd_ngroups, A.pd_cap);
  }
  A.dict_rows = PD_NB * BR;
  A.dict_bytes = (int64_t)total16 * 16 + 4 * (int64_t)(ntiles + 1);
  A.dict_state = 1;
  return true;
}
// build (or decline) the tile-local pattern/value dictionaries of A; called lazily from the first un-captured TMA SpMV.
// Column patterns relative to the block row (square stencil operators) are tried first; rectangular operators whose
// patterns do not repeat that way (interpolation, restriction) get the explicit-first-column mode.
template <int BR, int BC>
static void build_pd(const Csr &A, int force) {
  if (A.nrows == A.ncols * BR / BC && build_pd_mode<BR, BC>(A, force, 0, false)) return; // same node space for rows and columns
  if (build_pd_mode<BR, BC>(A, force, 1, false)) return;
  if (A.nrows != A.ncols * BR / BC) build_pd_mode<BR, BC>(A, force, 0, false);
}
static void build_value_dict(const Csr &A) {
  static const bool off = getenv("B200SP_NO_VALUE_DICT") && atoi(getenv("B200SP_NO_VALUE_DICT"));
  static const int force = getenv("B200SP_VALUE_DICT") ? atoi(getenv("B200SP_VALUE_DICT")) : 0;
  A.dict_state = -1;
  if (off || A.no_value_dict || A.nrows == 0 || A.nnz < 4096) return;
  const int br = A.bcol.p ? A.blk_r : 1, bc = A.bcol.p ? A.blk_c : 1;
  A.pd_br = A.pd_bc = 1;
  if (br == 2 && bc == 2) { A.pd_br = 2; A.pd_bc = 2; build_pd<2, 2>(A, force); }
  else if (br == 2 && bc == 1) { A.pd_br = 2; A.pd_bc = 1; build_pd<2, 1>(A, force); }
  else if (br == 1 && bc == 2) { A.pd_br = 1; A.pd_bc = 2; build_pd<1, 2>(A, force); }
  else build_pd<1, 1>(A, force); // scalar rows: also the 3 x 3 / 1 x 3 node-block matrices of the 3-D problem
}
void csr_drop_value_dict(Csr &A) {
  A.pd_blob.release(); A.pd_off.release();
  A.pd_gstart.release(); A.pd_goff.release(); A.h_pd_gstart.clear(); A.pd_ngroups = 0;
  A.dict_state = 0; A.pd_cap = 0; A.dict_rows = 0; A.dict_bytes = 0;
  if (A.wait_order_rows < 0) A.wait_order_rows = 0; // the group order belongs to the dropped format
}

// returns false when the matrix does not fit the shared-memory tiling (caller falls back)
int spmv_tma_tile_rows() {
  static const int env_R = getenv("B200SP_TMA_R") ? atoi(getenv("B200SP_TMA_R")) : 0;
  return env_R ? env_R : TMA_TILE_ROWS;
}

// co-resident CTAs per SM for `smem` bytes of dynamic shared memory and `threads` threads per CTA.  Every CTA also
// reserves 1 KB of the SM's 228 KB: ignoring that asked for 8 CTAs per SM where 7 fit, and the persistent grid ran a
// second, nearly empty wave (C block: 0.171 ms instead of 0.112 ms).
static int ctas_per_sm(size_t smem, int threads) {
  int n = (int)((size_t)(228 * 1024) / (smem + 1024));
  if (n < 1) n = 1;
  if (n * threads > 2048) n = 2048 / threads;
  if (n > 32) n = 32;
  return n;
}
// ... and the register file: the persistent grids are sized to exactly one wave, so the number of co-resident CTAs must be
// what the hardware will really schedule for THIS kernel (registers included), not a shared-memory estimate -- one CTA too
// many per SM and 1/8 of the grid runs as a second, nearly empty wave (measured: 0.195 instead of 0.133 ms on the A block
// when the stage size dropped just below the 8-CTA shared-memory threshold while 72 registers allow 7).
template <class K>
static int ctas_per_sm(K kernel, size_t smem, int threads) {
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem) != cudaSuccess || n < 1) {
    cudaGetLastError();
    n = 1;
  }
  return std::min(n, ctas_per_sm(smem, threads));
}

// SpMV through the tile-local pattern/value dictionaries.  Builds them lazily (never while a CUDA graph is being recorded);
// returns false when the matrix has none (declined, or not built yet during a capture) -- the caller then uses its plain kernel.
// Serves the short-row matrices of the TMA class and, as scalar rows of up to 96 entries, the long-row 3-D operators.
bool csr_spmv_pd(const Csr &A, const XSrc &xs, double *y, const SpmvEpi &epi) {
  Ctx *c = A.ctx;
  if (A.dict_state == 0) { // lazily, never while a CUDA graph is being recorded
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    B2_CUDA(cudaStreamIsCapturing(c->stream, &st));
    if (st == cudaStreamCaptureStatusNone) build_value_dict(A);
  }
  const int dbr = A.pd_br, dbc = A.pd_bc; // the block shape the blobs were built for (1 x 1 for shapes without a pd kernel, e.g. 3 x 3)
  if (A.dict_state == 1 && (dbc == 1 || (reinterpret_cast<uintptr_t>(xs.x) & 15) == 0)) { // one thread per block row
    static const int env_st = getenv("B200SP_PD_STAGES") ? atoi(getenv("B200SP_PD_STAGES")) : 0;
    const int pd_stages = env_st >= 2 && env_st <= TMA_MAX_STAGES ? env_st : 2;
    const size_t smem_d = (size_t)A.pd_cap * pd_stages + 8 * TMA_MAX_STAGES;
    if (!(c->attr_mask & 8u)) {
      B2_CUDA(cudaFuncSetAttribute(k_spmv_pd<2, 2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
      B2_CUDA(cudaFuncSetAttribute(k_spmv_pd<2, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
      B2_CUDA(cudaFuncSetAttribute(k_spmv_pd<2, 2, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
      B2_CUDA(cudaFuncSetAttribute(k_spmv_pd<2, 2, 9>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
      B2_CUDA(cudaFuncSetAttribute(k_spmv_pd<2, 1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
      B2_CUDA(cudaFuncSetAttribute(k_spmv_pd<1, 2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
      B2_CUDA(cudaFuncSetAttribute(k_spmv_pd<1, 1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
      c->attr_mask |= 8u;
    }
    static const int env_ub = getenv("B200SP_PD_UB") ? atoi(getenv("B200SP_PD_UB")) : 0; // blocks in flight per thread (tuning)
    int psm;
    if (dbr == 2 && dbc == 2) psm = env_ub == 5 ? ctas_per_sm(k_spmv_pd<2, 2, 5>, smem_d, PD_NB) : env_ub == 9 ? ctas_per_sm(k_spmv_pd<2, 2, 9>, smem_d, PD_NB)
                                  : env_ub == 2 ? ctas_per_sm(k_spmv_pd<2, 2, 2>, smem_d, PD_NB) : ctas_per_sm(k_spmv_pd<2, 2, 3>, smem_d, PD_NB);
    else if (dbr == 2 && dbc == 1) psm = ctas_per_sm(k_spmv_pd<2, 1, 3>, smem_d, PD_NB);
    else if (dbr == 1 && dbc == 2) psm = ctas_per_sm(k_spmv_pd<1, 2, 3>, smem_d, PD_NB);
    else psm = ctas_per_sm(k_spmv_pd<1, 1, 3>, smem_d, PD_NB);
    static const int env_psm = getenv("B200SP_TMA_CTAS") ? atoi(getenv("B200SP_TMA_CTAS")) : 0;
    if (env_psm && env_psm < psm) psm = env_psm;
    const int nbrows = A.nrows / dbr;
    const int ntd = A.pd_ngroups; // groups of consecutive tiles
    const int gridd = ntd < c->num_sms * psm ? ntd : c->num_sms * psm;
    const int *order = nullptr;
    int n_nowait = 0;
    if (xs.wait_flags) { ensure_wait_order_groups(A, PD_NB * dbr); if (A.wait_order_rows == -(PD_NB * dbr)) { order = A.wait_order.p; n_nowait = A.wait_n_nowait; } }
    auto al16 = [](const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    SpmvEpi e2 = epi;
    e2.vec2 = dbr == 2 && al16(y) && al16(epi.z) && al16(epi.pm1) && al16(epi.pk) && al16(epi.dinv);
#define B200SP_PD_LAUNCH(BR_, BC_, UB_) \
  k_spmv_pd<BR_, BC_, UB_><<<gridd, PD_NB, smem_d, c->stream>>>(nbrows, A.pd_gstart.p, ntd, order, A.pd_goff.p, A.pd_blob.p, xs, y, e2, A.pd_cap, pd_stages, n_nowait)
    if (dbr == 2 && dbc == 2 && env_ub == 5) B200SP_PD_LAUNCH(2, 2, 5);
    else if (dbr == 2 && dbc == 2 && env_ub == 9) B200SP_PD_LAUNCH(2, 2, 9);
    else if (dbr == 2 && dbc == 2 && env_ub == 2) B200SP_PD_LAUNCH(2, 2, 2);
    else if (dbr == 2 && dbc == 2) B200SP_PD_LAUNCH(2, 2, 3);
    else if (dbr == 2 && dbc == 1) B200SP_PD_LAUNCH(2, 1, 3);
    else if (dbr == 1 && dbc == 2) B200SP_PD_LAUNCH(1, 2, 3);
    else B200SP_PD_LAUNCH(1, 1, 3);
#undef B200SP_PD_LAUNCH
    check_launch("k_spmv_pd");
    return true;
  }
  return false;
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
  if (!(c->attr_mask & 4u)) {
    B2_CUDA(cudaFuncSetAttribute(k_spmv_tma<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    B2_CUDA(cudaFuncSetAttribute(k_spmv_tma<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    c->attr_mask |= 4u;
  }
  if (tile_list && R != TMA_TILE_ROWS) return false; // the lists were built for TMA_TILE_ROWS-row tiles
  const int ntiles = tile_list ? nlist : (A.nrows + R - 1) / R;
  if (ntiles <= 0) return true;
  const int per_sm = (env_U ? env_U == 6 : true) ? ctas_per_sm(k_spmv_tma<6>, smem, R) : ctas_per_sm(k_spmv_tma<3>, smem, R);
  int grid = ntiles < c->num_sms * per_sm ? ntiles : c->num_sms * per_sm;
  if (!tile_list && csr_spmv_pd(A, xs, y, epi)) return true; // tile-local dictionaries (built lazily)
  int n_nowait = 0;
  if (xs.wait_flags && !tile_list) { ensure_wait_order(A, R); if (A.wait_order_rows == R) { tile_list = A.wait_order.p; n_nowait = A.wait_n_nowait; } }
  if (A.bcol.p) { // block-compressed column index: 8 + 4/(BR*BC) bytes per nonzero (R is a multiple of BR)
    const int capb = ((cap / (A.blk_r * A.blk_c) + 8) + 3) & ~3;
    const size_t smem_b = ((size_t)cap * 8 + (size_t)capb * 4) * stages + 8 * TMA_MAX_STAGES;
    if (!(c->attr_mask & 16u)) {
      B2_CUDA(cudaFuncSetAttribute(k_spmv_tma_blk<2, 2, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
      B2_CUDA(cudaFuncSetAttribute(k_spmv_tma_blk<1, 2, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
      B2_CUDA(cudaFuncSetAttribute(k_spmv_tma_blk<2, 1, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
      c->attr_mask |= 16u;
    }
    const int psm = A.blk_r == 2 && A.blk_c == 2 ? ctas_per_sm(k_spmv_tma_blk<2, 2, 6>, smem_b, R)
                    : A.blk_r == 1 && A.blk_c == 2 ? ctas_per_sm(k_spmv_tma_blk<1, 2, 6>, smem_b, R) : ctas_per_sm(k_spmv_tma_blk<2, 1, 6>, smem_b, R);
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
