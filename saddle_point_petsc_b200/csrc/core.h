// core.h -- internal C++ declarations of libb200sp (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include "nccl_dyn.h"
#include <cstdint>
#include <cstdio>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/b200sp.h"

namespace b200sp {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

#define B2_CUDA(call)                                                                                   \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess)                                                                              \
      throw ::b200sp::Error(B200SP_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " at " + \
                                                 __FILE__ + ":" + std::to_string(__LINE__));            \
  } while (0)
#define B2_NCCL(call)                                                                                   \
  do {                                                                                                  \
    ncclResult_t e_ = (call);                                                                           \
    if (e_ != ncclSuccess)                                                                              \
      throw ::b200sp::Error(B200SP_ERR_NCCL, std::string(#call) + ": " + ::b200sp::nccl().GetErrorString(e_)); \
  } while (0)
#define B2_REQUIRE(cond, msg)                                                 \
  do {                                                                        \
    if (!(cond)) throw ::b200sp::Error(B200SP_ERR_ARG, std::string(msg));      \
  } while (0)

constexpr int RED_MAX_BLOCKS = 1024; // upper bound on the grid of a reduction kernel
constexpr int RED_MAX_OUT = 32;      // results per reduction launch
constexpr int N_SCALARS = 128;

struct ProfEntry { double ms = 0; int64_t n = 0; };

struct Ctx {
  int device = 0, rank = 0, size = 1;
  cudaStream_t stream = nullptr;  // compute stream (every kernel of the library)
  cudaStream_t stream2 = nullptr; // halo / copy stream
  ncclComm_t comm = nullptr;
  struct Comm *dcomm = nullptr;   // set when size > 1 (NCCL or in-process thread group), see dist.h
  struct Collective *reducer = nullptr; // Krylov reductions (peer-to-peer kernel when the communicator allows it)
  int num_sms = 148;
  double *d_partials = nullptr; // [RED_MAX_BLOCKS][RED_MAX_OUT]
  unsigned *d_ticket = nullptr;
  double *d_scalars = nullptr; // device result slots
  double *h_scalars = nullptr; // pinned mirror
  int *h_err = nullptr, *d_err = nullptr; // pinned+mapped error word: device-side waits that time out report here
  int64_t launches = 0;
  unsigned attr_mask = 0; // which kernel families had their MaxDynamicSharedMemorySize raised ON THIS CONTEXT'S DEVICE (the attribute is per device)
  bool profile = false;
  std::map<std::string, ProfEntry> prof;
  cudaEvent_t pev0 = nullptr, pev1 = nullptr, tev0 = nullptr, tev1 = nullptr;
  ~Ctx();
  void sync() { B2_CUDA(cudaStreamSynchronize(stream)); }
  // global sum of k doubles held in d (device) -> host array (blocking); NCCL all-reduce when size>1
  void fetch_scalars(const double *d, int k, double *host);
};

// bracket for one kernel launch: counts it, and in profile mode times it with events
struct LaunchScope {
  Ctx *c;
  const char *cls;
  LaunchScope(Ctx *ctx, const char *k) : c(ctx), cls(k) {
    c->launches++;
    if (c->profile) cudaEventRecord(c->pev0, c->stream);
  }
  ~LaunchScope() {
    if (c->profile) {
      cudaEventRecord(c->pev1, c->stream);
      cudaEventSynchronize(c->pev1);
      float ms = 0;
      cudaEventElapsedTime(&ms, c->pev0, c->pev1);
      auto &e = c->prof[cls];
      e.ms += ms;
      e.n++;
    }
  }
};
inline void check_launch(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw Error(B200SP_ERR_CUDA, std::string("launch ") + what + ": " + cudaGetErrorString(e));
}

template <class T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  DevBuf() {}
  explicit DevBuf(size_t n_) { alloc(n_); }
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DevBuf &operator=(DevBuf &&o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void alloc(size_t n_) {
    release();
    n = n_;
    if (cudaMalloc((void **)&p, (n ? n : 1) * sizeof(T)) != cudaSuccess) {
      p = nullptr;
      cudaGetLastError();
      throw Error(B200SP_ERR_MEM, "cudaMalloc of " + std::to_string(n * sizeof(T)) + " bytes failed");
    }
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  void zero(cudaStream_t s) { B2_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
};

// ------------------------------------------------------------------ Vec
struct Vec {
  Ctx *ctx;
  int64_t n;
  DevBuf<double> buf;
  double *d; // == buf.p, or a view into another allocation
  Vec(Ctx *c, int64_t n_) : ctx(c), n(n_), buf((size_t)n_ + 2), d(buf.p) { buf.zero(c->stream); }
  Vec(Ctx *c, double *view, int64_t n_) : ctx(c), n(n_), d(view) {}
};

// ------------------------------------------------------------------ DMDA
struct Layout;
struct Halo;
struct Dmda {
  Ctx *ctx = nullptr;
  int M = 0, N = 0;                       // nodes
  int pm = 1, pn = 1;                     // process grid
  int xs = 0, ys = 0, xm = 0, ym = 0;     // owned node box of this rank
  std::vector<int> lx, ly;
  std::shared_ptr<Layout> layout;         // null on a plain single-rank grid
  std::shared_ptr<Halo> halo;             // ghost ring of this rank (null on one rank)
};

// 3-D DMDA (BASELINE config 4): owned node box of this rank, ext-box column lookup table, ghost halo (null on one rank)
struct Dmda3 {
  Ctx *ctx = nullptr;
  int M = 0, N = 0, P = 0;
  int pm = 1, pn = 1, pp = 1;
  int xs = 0, ys = 0, zs = 0, xm = 0, ym = 0, zm = 0;
  int64_t g0 = 0;                   // first global (PETSc numbering) node id of this rank
  DevBuf<int> lut;                  // (xm+2)(ym+2)(zm+2): local column node id (owned: local index, ghost: n_owned + ghost index, -1 outside)
  std::shared_ptr<Halo> halo;
  std::shared_ptr<Layout> layout;   // carries size and rstart for the general halo
};

// ------------------------------------------------------------------ Mat
enum SpmvKernel { SPMV_STREAM = 0, SPMV_VECTOR = 1, SPMV_BLOCK = 2, SPMV_TMA = 3, SPMV_NODE = 4 }; // NODE: warp per node-block row (3-D blocks)
struct XSrc;
struct SpmvEpi;
constexpr int TMA_TILE_ROWS = 128; // rows per tile of the TMA SpMV kernel (profiles/r01_tma_tile_sweep.txt)
int spmv_tma_tile_rows();
bool csr_try_block_index(struct Csr &A, int br, int bc); // build the block-compressed column index if the pattern allows
void csr_drop_value_dict(struct Csr &A);                 // the values changed: forget the value dictionary (rebuilt lazily)
bool csr_spmv_tma(const struct Csr &A, const XSrc &xs, double *y, const SpmvEpi &epi, const int *tile_list = nullptr, int nlist = 0);
bool csr_spmv_pd(const struct Csr &A, const XSrc &xs, double *y, const SpmvEpi &epi); // tile-local dictionaries; false: matrix has none

struct Csr {
  Ctx *ctx = nullptr;
  int nrows = 0, ncols = 0;
  int64_t nnz = 0;
  DevBuf<int> rowptr;   // nrows+1
  DevBuf<int> col;      // nnz + PAD
  DevBuf<double> val;   // nnz + PAD
  // spmv plan
  int kernel = SPMV_VECTOR;
  int max_row_nnz = 0;
  int max_group_nnz = 0; // max nnz of any 32-row group (+1 alignment slack)
  int lanes_per_row = 32;
  int64_t hist[14] = {0};
  // grid metadata when the matrix came from DMDA assembly (for -pc_type mg); 0 = unknown
  int grid_M = 0, grid_N = 0, dof_r = 0, dof_c = 0;
  std::string tag = "spmv"; // profile class of this matrix's SpMV launches ("spmv:A", "spmv:Bt", ...)
  // row-partitioned (MPIAIJ-like) matrix: local rows; columns < ncols are owned (local ids), columns >= ncols are
  // ghost nodes in MPIAIJ garray order, read by the kernels from the halo buffer
  std::shared_ptr<Halo> halo;     // column-space halo (null on one rank); column ids >= ncols are ghost ids + ncols
  // block-compressed column index (kernels_spmv_tma.cu): one block-column id per blk_r x blk_c node block
  DevBuf<int> bptr, bcol;
  int blk_r = 1, blk_c = 1;
  // tile-local pattern/value dictionaries (kernels_spmv_tma.cu, "pd" format): one blob per tile of 128 block rows holding
  // the distinct values grouped by stencil position, the distinct column patterns and one byte per nonzero; built lazily
  // by the first SpMV (state 0 = not tried, 1 = in use, -1 = declined), dropped when values change
  mutable DevBuf<unsigned char> pd_blob;
  mutable DevBuf<int> pd_off;                 // blob offsets per tile, in 16-byte units
  mutable DevBuf<int> pd_gstart, pd_goff;     // first tile / blob offset (16-byte units) of every group of consecutive tiles (one group per pipeline stage)
  mutable std::vector<int> h_pd_gstart;
  mutable int dict_state = 0, pd_cap = 0, pd_ngroups = 0, dict_rows = 0; // pd_cap: largest group (bytes); dict_rows: rows per tile
  mutable int pd_br = 1, pd_bc = 1;           // block shape the blobs were built for
  mutable int64_t dict_bytes = 0;
  bool no_value_dict = false;                 // b200sp_mat_set_spmv_format: keep the plain value stream
  // tile order for kernels that wait for the halo themselves: tiles (of wait_order_rows rows) without ghost columns first
  mutable DevBuf<int> wait_order;
  mutable int wait_order_rows = 0, wait_n_nowait = 0;
  // TMA_TILE_ROWS-row tiles without / with ghost columns: the interior tiles are multiplied while the halo travels
  DevBuf<int> tiles_interior, tiles_boundary;
  int n_tiles_interior = 0, n_tiles_boundary = 0;
  int halo_dof = 0;               // dof per node of the column space
  std::shared_ptr<Layout> layout;  // node layout of a square DMDA matrix (multigrid coarsening)
  int64_t row_gstart = 0, col_gstart = 0; // first global row / owned global column of this rank (PETSc numbering)
  void plan();           // histogram + kernel choice (device reduction)
  int64_t state = 0;     // bumped whenever values change (PetscObjectState): a KSP set up on an older state sets up again
};
constexpr int CSR_PAD = 8; // zero entries appended to col/val so vector loads may overrun a row tile

struct Mat {
  Ctx *ctx = nullptr;
  bool nest = false;
  std::shared_ptr<Csr> csr;               // plain matrix
  std::shared_ptr<Csr> blk[2][2];         // nest blocks (blk[1][1] may be null)
  int64_t state() const { // sum of the block states
    if (!nest) return csr ? csr->state : 0;
    int64_t s = 0;
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) if (blk[i][j]) s += blk[i][j]->state;
    return s;
  }
  int nrows() const { return nest ? blk[0][0]->nrows + blk[1][0]->nrows : csr->nrows; }
  int ncols() const { return nest ? blk[0][0]->ncols + blk[0][1]->ncols : csr->ncols; }
};

// ------------------------------------------------------------------ kernels (launchers)
// vectors (kernels_vec.cu) -- all on ctx->stream
void vec_set(Ctx *c, int64_t n, double a, double *y);
void vec_copy(Ctx *c, int64_t n, const double *x, double *y);
void vec_scale(Ctx *c, int64_t n, double a, double *y);
void vec_axpy(Ctx *c, int64_t n, double a, const double *x, double *y);                  // y += a x
void vec_aypx(Ctx *c, int64_t n, double a, const double *x, double *y);                  // y = x + a y
void vec_waxpy(Ctx *c, int64_t n, double a, const double *x, const double *y, double *w); // w = a x + y
void vec_axpbypcz(Ctx *c, int64_t n, double a, const double *x, double b, const double *y, double cc, const double *z, double *w); // w = a x + b y + cc z
// w = a x + b y + cc (d .* z)   (Chebyshev update fused with the Jacobi application; d may be null)
void vec_cheb_update(Ctx *c, int64_t n, double a, const double *x, double b, const double *y, double cc, const double *d, const double *z, double *w);
void vec_pointwise_mult(Ctx *c, int64_t n, const double *x, const double *y, double *w);
void vec_reciprocal_safe(Ctx *c, int64_t n, double *d); // d = 1/(d==0?1:d)   (PCJACOBI setup)
void vec_hash(Ctx *c, int64_t n, double *v);
void vec_hash_natural(Ctx *c, int xs, int ys, int xm, int ym, int M, int dof, double *v); // same values as vec_hash on one rank, by natural index
void vec_scatter_set(Ctx *c, int64_t n, const int *idx, double val, double *y); // y[idx[i]] = val
void vec_permute_scatter(Ctx *c, int64_t n, const int *map, const double *in, double *out); // out[map[i]] = in[i] where map[i] >= 0
void vec_permute_gather(Ctx *c, int64_t n, const int *map, const double *in, double *out);  // out[i] = in[map[i]]
// reductions: results land in device slot `out` (k doubles); fetch with ctx->fetch_scalars
void vec_dot(Ctx *c, int64_t n, const double *x, const double *y, double *out);
void vec_mdot(Ctx *c, int64_t n, int k, const double *w, const double *V, int64_t ld, double *out); // out[j] = w . V_j
// w -= sum_j h[j] V_j (h on device), out[0] = ||w_new||^2
void vec_maxpy_norm2(Ctx *c, int64_t n, int k, double *w, const double *V, int64_t ld, const double *h_dev, double *out);
void vec_maxpy(Ctx *c, int64_t n, int k, double *w, const double *V, int64_t ld, const double *coef_dev); // w += sum coef[j] V_j
// y = x * (1/sqrt(*nrm2_dev))  (normalisation without a host round trip)
void vec_scale_inv_sqrt(Ctx *c, int64_t n, const double *nrm2_dev, const double *x, double *y);

// spmv (kernels_spmv.cu)
// where the kernels read x from: owned columns from the caller's vector, ghost columns (>= n_owned) from the halo buffer
struct XSrc {
  const double *x, *ghost;
  int n_owned;
  const unsigned long long *seq = nullptr; // peer-to-peer halos: exchanges started so far; the latest one is in parity (seq-1)&1
  long long ghost_stride = 0;
  // when set, the kernel itself waits for the neighbours' pushes (one lane per incoming message polls this rank's
  // flag words) before its first ghost read, instead of a separate wait kernel
  const unsigned long long *wait_flags = nullptr;
  int wait_nmsg = 0;
  int *wait_err = nullptr;
#ifdef __CUDACC__
  __device__ __forceinline__ double load(int c) const {
    if (c < n_owned) return __ldg(x + c);
    const double *g = ghost + (c - n_owned);
    if (seq) g += ((*seq - 1ull) & 1ull) * ghost_stride;
    return *g; // plain load: the buffer is written by a peer GPU (never through the non-coherent path)
  }
  __device__ __forceinline__ double2 load2(int c) const { // columns c, c+1 of one node (c even, x 16-byte aligned)
    if (c < n_owned) return __ldg(reinterpret_cast<const double2 *>(x + c));
    const double *g = ghost + (c - n_owned);
    if (seq) g += ((*seq - 1ull) & 1ull) * ghost_stride;
    return make_double2(g[0], g[1]);
  }
#endif
};
// ---- halo push fused into the producing SpMV (peer-to-peer halos) ----------------------------------------------
// When the caller knows that the vector an SpMV writes is the next one to be multiplied by a row-partitioned matrix,
// the kernel itself stores the values its neighbours need straight into their ghost buffers (NVLink) as it produces
// them, and the last CTA to finish raises the neighbours' sequence flags -- the separate push kernel, its launch and
// its dependency on the whole producer disappear.  Protocol (parity, flags, sequence counter) as in dist.h.
struct HaloP2PMsg { double *peer_ghost; unsigned long long *peer_flag; long long peer_stride; int send_off, send_cnt, peer_recv_off, pad; };
struct PushOut {
  const unsigned char *grp = nullptr; // per 64 owned nodes: does any of them go to a neighbour?  (null: no push)
  const int *node_ent = nullptr;      // per owned node: (first entry << 3) | number of entries (<= 7: a corner node of a 3-D box has 7 neighbours), 0 = none
  const int2 *ents = nullptr;         // {message, position of the node inside that message}
  const HaloP2PMsg *msgs = nullptr;
  unsigned long long *seq = nullptr;
  unsigned *ticket = nullptr;
  int nmsg = 0, dof = 1;
#ifdef __CUDACC__
  __device__ __forceinline__ void row(int r, double v) const { // row r of the produced vector has the value v
    const int node = dof == 2 ? r >> 1 : (dof == 1 ? r : r / dof), c = r - node * dof;
    if (!grp[node >> 6]) return;
    const int e = node_ent[node];
    if (!e) return;
    const unsigned long long par = *seq & 1ull;
    for (int k = 0; k < (e & 7); ++k) {
      const int2 t = ents[(e >> 3) + k];
      const HaloP2PMsg &g = msgs[t.x];
      g.peer_ghost[par * g.peer_stride + (long long)(g.peer_recv_off + t.y) * dof + c] = v;
    }
  }
  __device__ __forceinline__ void node2(int node, double v0, double v1) const { // both dofs of a node (dof == 2)
    if (!grp[node >> 6]) return;
    const int e = node_ent[node];
    if (!e) return;
    const unsigned long long par = *seq & 1ull;
    for (int k = 0; k < (e & 7); ++k) {
      const int2 t = ents[(e >> 3) + k];
      const HaloP2PMsg &g = msgs[t.x];
      *reinterpret_cast<double2 *>(g.peer_ghost + par * g.peer_stride + (long long)(g.peer_recv_off + t.y) * 2) = make_double2(v0, v1);
    }
  }
  // called by EVERY thread of EVERY CTA once, after its last row: one system-scope fence per CTA, a ticket, and the
  // last CTA raises the flags in the neighbours' memory and advances the sequence counter
  __device__ __forceinline__ void finish() const {
    __shared__ bool s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence_system();
      s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence_system();
    const unsigned long long k = *seq;
    __syncthreads(); // everyone has read seq before thread 0 advances it
    if ((int)threadIdx.x < nmsg) {
      volatile unsigned long long *f = msgs[threadIdx.x].peer_flag;
      *f = k + 1ull;
    }
    if (threadIdx.x == 0) { *ticket = 0u; *seq = k + 1ull; }
  }
#endif
};

// fused epilogue of every SpMV kernel, applied to the row sum s of row r:
//   cheb == 0:  beta_z*z[r] + alpha*s                                   (z may be null; z may alias y)
//   cheb == 1:  ca*pm1[r] + cb*pk[r] + cc*(dinv[r]*(z[r] - s))          (z = b; dinv may be null)  -- the same
//               operations, in the same order, as residual + PCJACOBI + VecAXPBYPCZ done separately
struct SpmvEpi {
  double alpha = 1.0;
  const double *z = nullptr;
  double beta_z = 0.0;
  int cheb = 0;
  const double *pm1 = nullptr, *pk = nullptr, *dinv = nullptr;
  double ca = 0.0, cb = 0.0, cc = 0.0;
  int vec2 = 0; // set by the launcher of the block-row kernel, see apply2_store
  PushOut push; // fused halo push of the produced vector (grp == null: none)
#ifdef __CUDACC__
  __device__ __forceinline__ double apply(double s, int r) const {
    if (cheb) {
      double t = z[r] - s;
      if (dinv) t = t * dinv[r];
      return ca * pm1[r] + cb * pk[r] + cc * t;
    }
    double v = alpha * s;
    if (z) v = beta_z * z[r] + v;
    return v;
  }
  // rows r, r+1 (r even) of one node with 16-byte loads/stores; same per-element operations as apply().  Only when
  // vec2 is set (the launcher checked that y and every operand are 16-byte aligned).
  __device__ __forceinline__ void apply2_store(double s0, double s1, int r, double *y, double &v0, double &v1) const {
    if (cheb) {
      const double2 zz = *reinterpret_cast<const double2 *>(z + r);
      double t0 = zz.x - s0, t1 = zz.y - s1;
      if (dinv) { const double2 d = *reinterpret_cast<const double2 *>(dinv + r); t0 = t0 * d.x; t1 = t1 * d.y; }
      const double2 a = *reinterpret_cast<const double2 *>(pm1 + r), b = *reinterpret_cast<const double2 *>(pk + r);
      v0 = ca * a.x + cb * b.x + cc * t0;
      v1 = ca * a.y + cb * b.y + cc * t1;
    } else {
      v0 = alpha * s0; v1 = alpha * s1;
      if (z) { const double2 zz = *reinterpret_cast<const double2 *>(z + r); v0 = beta_z * zz.x + v0; v1 = beta_z * zz.y + v1; }
    }
    *reinterpret_cast<double2 *>(y + r) = make_double2(v0, v1);
  }
#endif
};
// push_to: the produced vector y is the next one multiplied by a matrix whose column halo is *push_to (dof values per
// node): push its boundary values to the neighbours now (fused into the kernel when possible); the consumer then only waits
void csr_spmv(const Csr &A, const double *x, double *y, double alpha = 1.0, const double *z = nullptr, double beta_z = 0.0, bool reuse_halo = false,
              Halo *push_to = nullptr, int push_dof = 0);
// reuse_halo: the ghost values of this x are already in the halo buffer (previous SpMV with the same x and halo)
void csr_spmv_epi(const Csr &A, const double *x, double *y, const SpmvEpi &epi, bool reuse_halo = false, Halo *push_to = nullptr, int push_dof = 0);
void csr_get_diagonal(const Csr &A, double *d);
void csr_zero_rows_cols(Csr &A, int n, const int *rows_host, double diag, bool do_rows, bool do_cols, bool set_diag);

// setup utilities (kernels_setup.cu)
void exclusive_scan_i32(Ctx *c, const int *in, int *out, int64_t n, int *total_host); // out[n] = total too (out has n+1)
std::shared_ptr<Csr> csr_from_host(Ctx *c, int nrows, int ncols, const int *rowptr, const int *col, const double *val);
std::shared_ptr<Csr> csr_from_coo_host(Ctx *c, int nrows, int ncols, int64_t ncoo, const int *row, const int *col, const double *val);
std::shared_ptr<Csr> csr_transpose(const Csr &A);
std::shared_ptr<Csr> csr_matmat(const Csr &A, const Csr &B);
std::shared_ptr<Csr> csr_matmat_dist(const Csr &A, const Csr &B); // row-partitioned operands (dist_spgemm.cu), collective
std::shared_ptr<Csr> spgemm_raw(Ctx *c, int nrowsA, const int *rpa, const int *ca, const double *va, const int *rpb, const int *cb, const double *vb, int ncolsC);
void csr_copy_distribution(Csr &C, const Csr &like);
std::shared_ptr<Csr> csr_extract_fields(const Csr &A, int bs, const std::vector<int> &split_of_field, int rs, int cs); // MatCreateSubMatrix, strided ISs
std::shared_ptr<Csr> csr_scale_cols(const Csr &A, const double *d);            // A * diag(d)
std::shared_ptr<Csr> csr_scale_rows(const Csr &A, const double *d);            // diag(d) * A
// aggregation multigrid set-up (kernels_amg.cu; the algorithm oracle/sp_oracle_amg.c defines)
int amg_aggregate(const Csr &A, int bs, double theta, int order, DevBuf<int> &agg);        // agg[node] = aggregate id or -1; returns the number of aggregates
std::shared_ptr<Csr> amg_tentative(Ctx *c, int nn, int bs, const DevBuf<int> &agg, int nagg, const int *w, DevBuf<int> &wc);
std::shared_ptr<Csr> amg_smooth_prolongator(const Csr &A, const Csr &Pt, double omega); // Pt - omega D^-1 A Pt
std::shared_ptr<Csr> csr_add_scaled(const Csr &A, double a, const Csr &B);     // A + a B (same pattern required)
void dense_inverse_from_csr(const Csr &A, double *Ainv); // n x n row-major, device
void dense_matvec(Ctx *c, int n, const double *Ainv, const double *x, double *y);

// assembly (kernels_assembly.cu)
std::shared_ptr<Csr> assemble_stress(const Dmda &da, int as_written, int coeff_kind = 0);
void assemble_rhs(const Dmda &da, int as_written, int kind, double *f);
void assemble_kkt(const Dmda &da, std::shared_ptr<Csr> *Bt, std::shared_ptr<Csr> *B, std::shared_ptr<Csr> *C, std::shared_ptr<Csr> *Q);
void assemble_constraints(const Dmda &da, std::shared_ptr<Csr> *B, std::shared_ptr<Csr> *Bt); // the reference's 4 dense constraint rows
// 3-D assembly (kernels_assembly3d.cu)
void dmda3_build_lut(Dmda3 &da, const std::vector<int> &host_lut);
std::shared_ptr<Csr> assemble3_stress(const Dmda3 &da);
void assemble3_rhs(const Dmda3 &da, int kind, double *f);
void assemble3_kkt(const Dmda3 &da, std::shared_ptr<Csr> *Bt, std::shared_ptr<Csr> *B, std::shared_ptr<Csr> *C, std::shared_ptr<Csr> *Q);
std::vector<int> dmda3_bc_ids(const Dmda3 &da, int dof);
std::shared_ptr<Csr> interp_q1(Ctx *c, int Mc, int Nc, int dof, int bc);
std::shared_ptr<Csr> restrict_q1(Ctx *c, int Mc, int Nc, int dof, int bc); // = interp_q1^T, built directly
std::shared_ptr<Csr> interp_q1_dist(const Dmda &fine, const Dmda &coarse, int dof, int bc);
std::shared_ptr<Csr> restrict_q1_dist(const Dmda &fine, const Dmda &coarse, int dof, int bc);
std::shared_ptr<Csr> csr_alloc_public(Ctx *c, int nrows, int ncols, int64_t nnz);
std::vector<int> dmda_bc_ids(const Dmda &da, int dof);

// host-only DMDA index arithmetic (dmda.cpp part of capi)
void dmda_proc_grid(int M, int N, int size, int *m, int *n);
void dmda_ownership(int M, int m, int *lx);

} // namespace b200sp
