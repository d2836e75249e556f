// kernels_assembly.cu -- device assembly of the reference's Q1 discretisation (replaces
// AssembleOperator_Laplace / AssembleRHS_Laplace / ApplyBC_Laplace and the stubbed
// AssembleOperator_Constraints, src/Discretization.c:130-290, plus DMCreateMatrix, SaddlePointProblem.c:42).
//
// Two phases, both bit-exact against MatSetValuesStencil(ADD_VALUES) + MatAssemblyEnd on one rank:
//   1. element kernel: one thread per element evaluates the element matrices in the reference's exact
//      operation order (no FMA contraction: this file is compiled with -fmad=false) and stores them
//      entry-major (SoA) so stores and the later gathers are coalesced.  This is the COO stage: the
//      (row, col) of every value is implied by index arithmetic (DMDAGetElementEqnums, :377-395).
//   2. CSR stage: one thread per matrix row writes the DMCreateMatrix box-stencil pattern (ascending
//      columns, explicit zeros kept) and sums, for every entry, the <= 4 element contributions in the
//      reference's element order (j outer, i inner, :146-147) starting from +0.0.
// The reference's truncated Gauss abscissa 0.57735026919 (:52-55) is kept on purpose.
#include "dev.cuh"
#include "dist.h"
#include <algorithm>

namespace b200sp {

std::shared_ptr<Csr> csr_alloc_public(Ctx *c, int nrows, int ncols, int64_t nnz);

namespace {

struct ElemBox { int ex0, ey0, enx, eny; }; // element range held in the SoA arrays

__device__ __forceinline__ void gauss_point(int p, double xi[2]) {
  // ConstructGaussQuadratureQ12D, Discretization.c:49-63 (weights are 1.0)
  const double g = 0.57735026919;
  xi[0] = (p < 2) ? -g : g;
  xi[1] = (p == 0 || p == 3) ? -g : g;
}
__device__ __forceinline__ void q1_Ni(const double xi_[2], double Ni[4]) { // :65-76
  const double xi = xi_[0], eta = xi_[1];
  Ni[0] = 0.25 * (1.0 - xi) * (1.0 - eta);
  Ni[1] = 0.25 * (1.0 - xi) * (1.0 + eta);
  Ni[2] = 0.25 * (1.0 + xi) * (1.0 + eta);
  Ni[3] = 0.25 * (1.0 + xi) * (1.0 - eta);
}
__device__ __forceinline__ void q1_GNi(const double xi_[2], double GNi[2][4]) { // :78-94
  const double xi = xi_[0], eta = xi_[1];
  GNi[0][0] = -0.25 * (1.0 - eta);
  GNi[0][1] = -0.25 * (1.0 + eta);
  GNi[0][2] = 0.25 * (1.0 + eta);
  GNi[0][3] = 0.25 * (1.0 - eta);
  GNi[1][0] = -0.25 * (1.0 - xi);
  GNi[1][1] = 0.25 * (1.0 - xi);
  GNi[1][2] = 0.25 * (1.0 + xi);
  GNi[1][3] = -0.25 * (1.0 + xi);
}
__device__ __forceinline__ void q1_GNx(const double GNi[2][4], const double *ec, double GNx[2][4], double *detJ) { // :96-128
  double Jac[2][2], invJ[2][2];
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < 4; ++i) s += GNi[c][i] * ec[i * 2 + d];
      Jac[c][d] = s;
    }
  const double J = Jac[0][0] * Jac[1][1] - Jac[0][1] * Jac[1][0];
  invJ[0][0] = Jac[1][1] / J;
  invJ[0][1] = -Jac[0][1] / J;
  invJ[1][0] = -Jac[1][0] / J;
  invJ[1][1] = Jac[0][0] / J;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    GNx[0][i] = invJ[0][0] * GNi[0][i] + invJ[0][1] * GNi[1][i];
    GNx[1][i] = invJ[1][0] * GNi[0][i] + invJ[1][1] * GNi[1][i];
  }
  *detJ = J;
}
// DMDASetUniformCoordinates + GetElementCoords (:25, :31-46)
__device__ __forceinline__ void element_coords(int M, int N, int ei, int ej, int as_written, double ec[8]) {
  const double hx = (1.0 - 0.0) / (double)(M - 1), hy = (1.0 - 0.0) / (double)(N - 1);
  const int di[4] = {0, 0, 1, 1}, dj[4] = {0, 1, 1, 0};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = ei + (as_written ? 0 : di[k]), j = ej + (as_written ? 0 : dj[k]);
    ec[2 * k + 0] = 0.0 + hx * (double)i;
    ec[2 * k + 1] = 0.0 + hy * (double)j;
  }
}

enum { WANT_K = 1, WANT_F = 2, WANT_KKT = 4, WANT_COEFF = 8 }; // COEFF: variable coefficient 1 + x(1-y)/2 at the Gauss point

// one thread per element; outputs entry-major: X[entry * nel + e]
__global__ void __launch_bounds__(128) k_elements(int M, int N, ElemBox eb, int as_written, int rhs_kind, int want,
                                                  double *__restrict__ Ke, double *__restrict__ Fe, double *__restrict__ Ge,
                                                  double *__restrict__ Ce, double *__restrict__ Qe) {
  const int64_t nel = (int64_t)eb.enx * eb.eny;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nel) return;
  const int ei = eb.ex0 + (int)(e % eb.enx), ej = eb.ey0 + (int)(e / eb.enx);
  double ec[8];
  element_coords(M, N, ei, ej, as_written, ec);

  if (want & WANT_K) { // FormStressOperatorQ12D, :293-332, coeff = 1 (:156-157)
    double K[64];
#pragma unroll
    for (int t = 0; t < 64; ++t) K[t] = 0.0;
    for (int p = 0; p < 4; ++p) {
      double xi[2], GNi[2][4], GNx[2][4], detJ, B[3][8], tD[3];
      gauss_point(p, xi);
      q1_GNi(xi, GNi);
      q1_GNx(GNi, ec, GNx, &detJ);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        B[0][2 * i] = GNx[0][i]; B[0][2 * i + 1] = 0.0;
        B[1][2 * i] = 0.0;       B[1][2 * i + 1] = GNx[1][i];
        B[2][2 * i] = GNx[1][i]; B[2][2 * i + 1] = GNx[0][i];
      }
      double coeff = 1.0; // an input of FormStressOperatorQ12D, one value per Gauss point (:151-157 sets 1.0)
      if (want & WANT_COEFF) {
        double Ni[4], xp = 0.0, yp = 0.0;
        q1_Ni(xi, Ni);
#pragma unroll
        for (int i = 0; i < 4; ++i) { xp += Ni[i] * ec[2 * i]; yp += Ni[i] * ec[2 * i + 1]; }
        coeff = 1.0 + 0.5 * xp * (1.0 - yp);
      }
      const double w = 1.0;
      tD[0] = 2.0 * w * detJ * coeff;
      tD[1] = 2.0 * w * detJ * coeff;
      tD[2] = w * detJ * coeff;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
          for (int k = 0; k < 3; ++k) K[i + 8 * j] += B[k][i] * tD[k] * B[k][j];
    }
#pragma unroll
    for (int t = 0; t < 64; ++t) Ke[(size_t)t * nel + e] = K[t];
  }
  if (want & WANT_F) { // FormLaplaceRHSQ12D + FormRHS, :334-374, 397-402
    double F[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) F[t] = 0.0;
    for (int p = 0; p < 4; ++p) {
      double xi[2], Ni[4], GNi[2][4], GNx[2][4], detJ, f_p[2];
      gauss_point(p, xi);
      q1_Ni(xi, Ni);
      q1_GNi(xi, GNi);
      q1_GNx(GNi, ec, GNx, &detJ);
      const double fac = 1.0 * detJ;
      if (rhs_kind == 0) { f_p[0] = 1.0; f_p[1] = 2.0; }
      else {
        double xp = 0.0, yp = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) { xp += Ni[i] * ec[2 * i]; yp += Ni[i] * ec[2 * i + 1]; }
        f_p[0] = 2.0 * yp - 1.0;
        f_p[1] = 1.0 - 2.0 * xp;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 2; ++c) F[i * 2 + c] += fac * Ni[i] * f_p[c];
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) Fe[(size_t)t * nel + e] = F[t];
  }
  if (want & WANT_KKT) { // KKT element blocks (ours; ex43 lineage named at src/main.c:1), see oracle or_element_kkt
    double G[32], C[16], Q[16];
#pragma unroll
    for (int t = 0; t < 32; ++t) G[t] = 0.0;
#pragma unroll
    for (int t = 0; t < 16; ++t) { C[t] = 0.0; Q[t] = 0.0; }
    for (int p = 0; p < 4; ++p) {
      double xi[2], Ni[4], GNi[2][4], GNx[2][4], detJ;
      gauss_point(p, xi);
      q1_Ni(xi, Ni);
      q1_GNi(xi, GNi);
      q1_GNx(GNi, ec, GNx, &detJ);
      const double fac = 1.0 * detJ;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int d = 0; d < 2; ++d)
#pragma unroll
          for (int j = 0; j < 4; ++j) G[(2 * i + d) * 4 + j] -= fac * GNx[d][i] * Ni[j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          C[i * 4 + j] -= fac * (Ni[i] * Ni[j] - 0.0625);
          Q[i * 4 + j] -= fac * (Ni[i] * Ni[j]);
        }
    }
#pragma unroll
    for (int t = 0; t < 32; ++t) Ge[(size_t)t * nel + e] = G[t];
#pragma unroll
    for (int t = 0; t < 16; ++t) { Ce[(size_t)t * nel + e] = C[t]; Qe[(size_t)t * nel + e] = Q[t]; }
  }
}

// local node number of node (i,j) inside element (ei,ej): DMDAGetElementEqnums order (:377-395)
__device__ __forceinline__ int local_node(int i, int j, int ei, int ej) {
  const int dx = i - ei, dy = j - ej;
  return dx == 0 ? (dy == 0 ? 0 : 1) : (dy == 0 ? 3 : 2);
}

struct GridBox { int M, N, xs, ys, xm, ym; }; // global node counts, owned node box

// row lengths of the box-stencil pattern (DMCreateMatrix): dofc * (clipped 3x3 box)
__global__ void __launch_bounds__(256) k_box_rowlen(GridBox g, int dofr, int dofc, int *len) {
  const int nrows = g.xm * g.ym * dofr;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x) {
    const int node = r / dofr;
    const int i = g.xs + node % g.xm, j = g.ys + node / g.xm;
    const int wx = 3 - (i == 0) - (i == g.M - 1), wy = 3 - (j == 0) - (j == g.N - 1);
    len[r] = dofc * wx * wy;
  }
}

// CSR stage.  E is the entry-major element array with `estride` = entries per element row:
//   transposed == 0: value(a, b) = E[(la * rowlen_e + lb)]  with la = ln_a*dofr + c, lb = ln_b*dofc + cc, rowlen_e = 4*dofc
//   transposed == 1: value(a, b) = E[(lb * (4*dofr) + la)]  (B = Ge^T read out of the gradient block)
__global__ void __launch_bounds__(256) k_box_fill(GridBox g, ElemBox eb, int dofr, int dofc, int transposed, const double *__restrict__ E,
                                                  const int *__restrict__ rowptr, int *__restrict__ col, double *__restrict__ val) {
  const int nrows = g.xm * g.ym * dofr;
  const int64_t nel = (int64_t)eb.enx * eb.eny;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x) {
    const int node = r / dofr, c = r % dofr;
    const int i = g.xs + node % g.xm, j = g.ys + node / g.xm;
    int p = rowptr[r];
    for (int jj = max(j - 1, 0); jj <= min(j + 1, g.N - 1); ++jj)
      for (int ii = max(i - 1, 0); ii <= min(i + 1, g.M - 1); ++ii) {
        // single-rank layout: local column node id == natural id inside the owned box
        const int cnode = (jj - g.ys) * g.xm + (ii - g.xs);
        const int ej0 = max(max(j, jj) - 1, 0), ej1 = min(min(j, jj), g.N - 2);
        const int ei0 = max(max(i, ii) - 1, 0), ei1 = min(min(i, ii), g.M - 2);
        for (int cc = 0; cc < dofc; ++cc) {
          double acc = 0.0;
          for (int ej = ej0; ej <= ej1; ++ej)
            for (int ei = ei0; ei <= ei1; ++ei) {
              const int la = local_node(i, j, ei, ej) * dofr + c;
              const int lb = local_node(ii, jj, ei, ej) * dofc + cc;
              const int entry = transposed ? lb * (4 * dofr) + la : la * (4 * dofc) + lb;
              const int64_t e = (int64_t)(ej - eb.ey0) * eb.enx + (ei - eb.ex0);
              acc += E[(size_t)entry * nel + e];
            }
          col[p] = cnode * dofc + cc;
          val[p] = acc;
          ++p;
        }
      }
  }
}

// ------------------------------------------------------------------------------------------------
// Distributed (row-partitioned) CSR stage.  A "row enumerator" lists the entries of one owned row in ascending
// (column node j, column node i, column dof) order; the builder classifies every column node as owned (diagonal
// block, local ids) or ghost (off-diagonal block, ghost ids in MPIAIJ garray order) exactly like MatMPIAIJ.
// Values are the same sums in the same order as on one rank: the assembled matrix is partition-independent.
struct RowBox { int xs, ys, xm, ym; };

struct BoxEnum { // A, B^T, B, C, Q: box stencil on one grid
  int M, N;
  RowBox rb;
  ElemBox eb;
  int dofr, dofc, transposed;
  const double *E;
  __device__ int nrows() const { return rb.xm * rb.ym * dofr; }
  template <class Emit> __device__ void enumerate(int r, Emit &emit) const {
    const int64_t nel = (int64_t)eb.enx * eb.eny;
    const int node = r / dofr, c = r % dofr;
    const int i = rb.xs + node % rb.xm, j = rb.ys + node / rb.xm;
    for (int jj = max(j - 1, 0); jj <= min(j + 1, N - 1); ++jj)
      for (int ii = max(i - 1, 0); ii <= min(i + 1, M - 1); ++ii) {
        const int ej0 = max(max(j, jj) - 1, 0), ej1 = min(min(j, jj), N - 2);
        const int ei0 = max(max(i, ii) - 1, 0), ei1 = min(min(i, ii), M - 2);
        for (int cc = 0; cc < dofc; ++cc) {
          double acc = 0.0;
          if (E)
            for (int ej = ej0; ej <= ej1; ++ej)
              for (int ei = ei0; ei <= ei1; ++ei) {
                const int la = local_node(i, j, ei, ej) * dofr + c;
                const int lb = local_node(ii, jj, ei, ej) * dofc + cc;
                const int entry = transposed ? lb * (4 * dofr) + la : la * (4 * dofc) + lb;
                const int64_t e = (int64_t)(ej - eb.ey0) * eb.enx + (ei - eb.ex0);
                acc += E[(size_t)entry * nel + e];
              }
          emit(ii, jj, cc, acc);
        }
      }
  }
};
struct InterpEnum { // P: rows = owned fine nodes, columns = coarse nodes
  int Mc, Nc, dof, bc;
  RowBox rb; // fine
  __device__ int nrows() const { return rb.xm * rb.ym * dof; }
  template <class Emit> __device__ void enumerate(int r, Emit &emit) const {
    const int Mf = 2 * Mc - 1, Nf = 2 * Nc - 1;
    const int node = r / dof, c = r % dof;
    const int i = rb.xs + node % rb.xm, j = rb.ys + node / rb.xm;
    const int fb = (i == 0 || i == Mf - 1 || j == 0 || j == Nf - 1);
    const int ni = (i & 1) ? 2 : 1, nj = (j & 1) ? 2 : 1;
    for (int b = 0; b < nj; ++b)
      for (int a = 0; a < ni; ++a) {
        const int ic = i / 2 + a, jc = j / 2 + b;
        const int cb = (ic == 0 || ic == Mc - 1 || jc == 0 || jc == Nc - 1);
        double w = (ni == 2 ? 0.5 : 1.0) * (nj == 2 ? 0.5 : 1.0);
        if (bc && (fb || cb)) w = 0.0;
        emit(ic, jc, c, w);
      }
  }
};
struct RestrictEnum { // R = P^T: rows = owned coarse nodes, columns = fine nodes
  int Mc, Nc, dof, bc;
  RowBox rb; // coarse
  __device__ int nrows() const { return rb.xm * rb.ym * dof; }
  template <class Emit> __device__ void enumerate(int r, Emit &emit) const {
    const int Mf = 2 * Mc - 1, Nf = 2 * Nc - 1;
    const int node = r / dof, c = r % dof;
    const int ic = rb.xs + node % rb.xm, jc = rb.ys + node / rb.xm;
    const int cb = (ic == 0 || ic == Mc - 1 || jc == 0 || jc == Nc - 1);
    for (int j = max(2 * jc - 1, 0); j <= min(2 * jc + 1, Nf - 1); ++j)
      for (int i = max(2 * ic - 1, 0); i <= min(2 * ic + 1, Mf - 1); ++i) {
        const int fb = (i == 0 || i == Mf - 1 || j == 0 || j == Nf - 1);
        double w = (i == 2 * ic ? 1.0 : 0.5) * (j == 2 * jc ? 1.0 : 0.5);
        if (bc && (fb || cb)) w = 0.0;
        emit(i, j, c, w);
      }
  }
};

struct CountEmit {
  int n = 0;
  __device__ void operator()(int, int, int, double) { ++n; }
};
struct FillEmit {
  ColSpace cs;
  int dofc, n_owned_cols, p;
  int *col; double *val;
  __device__ void operator()(int ci, int cj, int cc, double v) {
    if (ci >= cs.xs && ci < cs.xs + cs.xm && cj >= cs.ys && cj < cs.ys + cs.ym)
      col[p] = ((cj - cs.ys) * cs.xm + (ci - cs.xs)) * dofc + cc;
    else
      col[p] = n_owned_cols + cs.ring2ghost[cs.ring_id(ci, cj)] * dofc + cc; // ghost column
    val[p++] = v;
  }
};
template <class E> __global__ void __launch_bounds__(128) k_dist_count(E e, int *len) {
  const int nrows = e.nrows();
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x) {
    CountEmit em;
    e.enumerate(r, em);
    len[r] = em.n;
  }
}
template <class E> __global__ void __launch_bounds__(128) k_dist_fill(E e, ColSpace cs, int dofc, int n_owned_cols, const int *__restrict__ rp, int *col, double *val) {
  const int nrows = e.nrows();
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x) {
    FillEmit em{cs, dofc, n_owned_cols, rp[r], col, val};
    e.enumerate(r, em);
  }
}

__global__ void __launch_bounds__(256) k_tile_has_ghost(int nrows, int n_owned_cols, const int *__restrict__ rowptr, const int *__restrict__ col, int *tile_flag) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x) {
    bool g = false;
    for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) g |= col[k] >= n_owned_cols;
    if (g) tile_flag[r / TMA_TILE_ROWS] = 1;
  }
}

template <class E>
std::shared_ptr<Csr> build_dist(Ctx *c, const E &e, int nrows, const Halo &colh, int dofc, const char *tag) {
  ColSpace cs{colh.xs, colh.ys, colh.xm, colh.ym, colh.d_ring2ghost.p};
  const int ncols = colh.n_owned * dofc;
  DevBuf<int> len((size_t)nrows + 1), rp((size_t)nrows + 1);
  const int grid = std::max(1, std::min((nrows + 127) / 128, c->num_sms * 16));
  {
    LaunchScope ls(c, "assembly");
    k_dist_count<E><<<grid, 128, 0, c->stream>>>(e, len.p);
    check_launch("k_dist_count");
  }
  int nnz = 0;
  exclusive_scan_i32(c, len.p, rp.p, nrows, &nnz);
  auto A = csr_alloc_public(c, nrows, ncols, nnz);
  B2_CUDA(cudaMemcpyAsync(A->rowptr.p, rp.p, sizeof(int) * ((size_t)nrows + 1), cudaMemcpyDeviceToDevice, c->stream));
  {
    LaunchScope ls(c, "assembly");
    k_dist_fill<E><<<grid, 128, 0, c->stream>>>(e, cs, dofc, ncols, A->rowptr.p, A->col.p, A->val.p);
    check_launch("k_dist_fill");
  }
  c->sync();
  A->tag = tag;
  A->plan();
  // interior / boundary tile lists (TMA_TILE_ROWS rows per tile): a tile is "boundary" if any of its rows has a ghost column
  const int ntiles = (nrows + TMA_TILE_ROWS - 1) / TMA_TILE_ROWS;
  DevBuf<int> flag((size_t)ntiles + 1);
  flag.zero(c->stream);
  {
    LaunchScope ls(c, "assembly");
    k_tile_has_ghost<<<std::max(1, std::min((nrows + 255) / 256, c->num_sms * 16)), 256, 0, c->stream>>>(nrows, ncols, A->rowptr.p, A->col.p, flag.p);
    check_launch("k_tile_has_ghost");
  }
  std::vector<int> h((size_t)ntiles + 1), ti, tb;
  B2_CUDA(cudaMemcpyAsync(h.data(), flag.p, sizeof(int) * (size_t)ntiles, cudaMemcpyDeviceToHost, c->stream));
  c->sync();
  for (int t = 0; t < ntiles; ++t) (h[(size_t)t] ? tb : ti).push_back(t);
  A->n_tiles_interior = (int)ti.size();
  A->n_tiles_boundary = (int)tb.size();
  A->tiles_interior.alloc(ti.size() + 1);
  A->tiles_boundary.alloc(tb.size() + 1);
  if (!ti.empty()) B2_CUDA(cudaMemcpyAsync(A->tiles_interior.p, ti.data(), sizeof(int) * ti.size(), cudaMemcpyHostToDevice, c->stream));
  if (!tb.empty()) B2_CUDA(cudaMemcpyAsync(A->tiles_boundary.p, tb.data(), sizeof(int) * tb.size(), cudaMemcpyHostToDevice, c->stream));
  c->sync();
  return A;
}

// AssembleRHS_Laplace (:196-219): f[node,c] = sum of Fe over the <= 4 elements in (ej, ei) order
__global__ void __launch_bounds__(256) k_rhs_gather(GridBox g, ElemBox eb, const double *__restrict__ Fe, double *__restrict__ f) {
  const int nrows = g.xm * g.ym * 2;
  const int64_t nel = (int64_t)eb.enx * eb.eny;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x) {
    const int node = r >> 1, c = r & 1;
    const int i = g.xs + node % g.xm, j = g.ys + node / g.xm;
    double acc = 0.0;
    for (int ej = max(j - 1, 0); ej <= min(j, g.N - 2); ++ej)
      for (int ei = max(i - 1, 0); ei <= min(i, g.M - 2); ++ei) {
        const int entry = local_node(i, j, ei, ej) * 2 + c;
        const int64_t e = (int64_t)(ej - eb.ey0) * eb.enx + (ei - eb.ex0);
        acc += Fe[(size_t)entry * nel + e];
      }
    f[r] = acc;
  }
}

// AssembleOperator_Constraints (a stub in the reference, Discretization.c:277-283; B is 4 x nCols, SaddlePointProblem.c:48-49):
// the four dense rows defined in oracle/sp_oracle.c or_element_constraints (barycentre x / y, dilation moment, rotation
// moment about the domain centre).  One thread per node: for each of its <= 4 elements, in the reference's element order,
// the element's contribution to this node's six entries is summed over the Gauss points from +0.0 and then added -- the
// same operations in the same order as the element-vector + ADD_VALUES loop of the oracle.
__global__ void __launch_bounds__(128) k_constraint_rows(int M, int N, int *__restrict__ col, double *__restrict__ val) {
  const int nn = M * N, n = 2 * nn;
  for (int node = blockIdx.x * blockDim.x + threadIdx.x; node < nn; node += gridDim.x * blockDim.x) {
    const int i = node % M, j = node / M;
    double t0 = 0.0, t1 = 0.0, t2x = 0.0, t2y = 0.0, t3x = 0.0, t3y = 0.0;
    for (int ej = max(j - 1, 0); ej <= min(j, N - 2); ++ej)
      for (int ei = max(i - 1, 0); ei <= min(i, M - 2); ++ei) {
        double ec[8];
        element_coords(M, N, ei, ej, 0, ec);
        const int a = local_node(i, j, ei, ej);
        double c0 = 0.0, c1 = 0.0, c2x = 0.0, c2y = 0.0, c3x = 0.0, c3y = 0.0;
        for (int p = 0; p < 4; ++p) {
          double xi[2], Ni[4], GNi[2][4], GNx[2][4], detJ;
          gauss_point(p, xi);
          q1_Ni(xi, Ni);
          q1_GNi(xi, GNi);
          q1_GNx(GNi, ec, GNx, &detJ);
          const double fac = 1.0 * detJ;
          double xp = 0.0, yp = 0.0;
#pragma unroll
          for (int q = 0; q < 4; ++q) { xp += Ni[q] * ec[2 * q]; yp += Ni[q] * ec[2 * q + 1]; }
          const double rx = xp - 0.5, ry = yp - 0.5;
          const double w = fac * Ni[a];
          c0 += w;
          c1 += w;
          c2x += w * rx;
          c2y += w * ry;
          c3x -= w * ry;
          c3y += w * rx;
        }
        t0 += c0; t1 += c1; t2x += c2x; t2y += c2y; t3x += c3x; t3y += c3y;
      }
    col[node] = 2 * node;                 val[node] = t0;
    col[nn + node] = 2 * node + 1;        val[nn + node] = t1;
    col[2 * nn + 2 * node] = 2 * node;    val[2 * nn + 2 * node] = t2x;
    col[2 * nn + 2 * node + 1] = 2 * node + 1; val[2 * nn + 2 * node + 1] = t2y;
    col[2 * nn + n + 2 * node] = 2 * node;     val[2 * nn + n + 2 * node] = t3x;
    col[2 * nn + n + 2 * node + 1] = 2 * node + 1; val[2 * nn + n + 2 * node + 1] = t3y;
  }
}

// Q1 interpolation (DMCreateInterpolation_DA_2D_Q1 weights), see oracle or_interp_q1
__global__ void __launch_bounds__(256) k_interp_len(int Mc, int Nc, int dof, int *len) {
  const int Mf = 2 * Mc - 1, Nf = 2 * Nc - 1, nrows = Mf * Nf * dof;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x) {
    const int node = r / dof, i = node % Mf, j = node / Mf;
    len[r] = ((i & 1) ? 2 : 1) * ((j & 1) ? 2 : 1);
  }
}
__global__ void __launch_bounds__(256) k_interp_fill(int Mc, int Nc, int dof, int bc, const int *__restrict__ rowptr, int *col, double *val) {
  const int Mf = 2 * Mc - 1, Nf = 2 * Nc - 1, nrows = Mf * Nf * dof;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x) {
    const int node = r / dof, c = r % dof, i = node % Mf, j = node / Mf;
    const int fb = (i == 0 || i == Mf - 1 || j == 0 || j == Nf - 1);
    const int ni = (i & 1) ? 2 : 1, nj = (j & 1) ? 2 : 1;
    int p = rowptr[r];
    for (int b = 0; b < nj; ++b)
      for (int a = 0; a < ni; ++a) {
        const int ic = i / 2 + a, jc = j / 2 + b;
        const int cb = (ic == 0 || ic == Mc - 1 || jc == 0 || jc == Nc - 1);
        double w = (ni == 2 ? 0.5 : 1.0) * (nj == 2 ? 0.5 : 1.0);
        if (bc && (fb || cb)) w = 0.0;
        col[p] = (jc * Mc + ic) * dof + c;
        val[p] = w;
        ++p;
      }
  }
}

// restriction R = P^T written directly (rows = coarse dofs, ascending fine columns), same bc zeroing as P
__global__ void __launch_bounds__(256) k_restrict_len(int Mc, int Nc, int dof, int *len) {
  const int Mf = 2 * Mc - 1, Nf = 2 * Nc - 1, nrows = Mc * Nc * dof;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x) {
    const int node = r / dof, ic = node % Mc, jc = node / Mc;
    const int i0 = max(2 * ic - 1, 0), i1 = min(2 * ic + 1, Mf - 1), j0 = max(2 * jc - 1, 0), j1 = min(2 * jc + 1, Nf - 1);
    len[r] = (i1 - i0 + 1) * (j1 - j0 + 1);
  }
}
__global__ void __launch_bounds__(256) k_restrict_fill(int Mc, int Nc, int dof, int bc, const int *__restrict__ rowptr, int *col, double *val) {
  const int Mf = 2 * Mc - 1, Nf = 2 * Nc - 1, nrows = Mc * Nc * dof;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x) {
    const int node = r / dof, c = r % dof, ic = node % Mc, jc = node / Mc;
    const int cb = (ic == 0 || ic == Mc - 1 || jc == 0 || jc == Nc - 1);
    const int i0 = max(2 * ic - 1, 0), i1 = min(2 * ic + 1, Mf - 1), j0 = max(2 * jc - 1, 0), j1 = min(2 * jc + 1, Nf - 1);
    int p = rowptr[r];
    for (int j = j0; j <= j1; ++j)
      for (int i = i0; i <= i1; ++i) {
        const int fb = (i == 0 || i == Mf - 1 || j == 0 || j == Nf - 1);
        double w = (i == 2 * ic ? 1.0 : 0.5) * (j == 2 * jc ? 1.0 : 0.5);
        if (bc && (fb || cb)) w = 0.0;
        col[p] = (j * Mf + i) * dof + c;
        val[p] = w;
        ++p;
      }
  }
}

inline int grid_for(Ctx *c, int64_t n) {
  int64_t g = (n + 255) / 256;
  int64_t cap = (int64_t)c->num_sms * 16;
  return (int)std::max<int64_t>(1, std::min(g, cap));
}

struct ElemArrays {
  ElemBox eb;
  DevBuf<double> Ke, Fe, Ge, Ce, Qe;
};

ElemBox element_box(const Dmda &da) {
  // ghost-element recomputation: every element touching an owned node -> no assembly communication
  ElemBox eb;
  eb.ex0 = std::max(da.xs - 1, 0);
  eb.ey0 = std::max(da.ys - 1, 0);
  int ex1 = std::min(da.xs + da.xm - 1, da.M - 2), ey1 = std::min(da.ys + da.ym - 1, da.N - 2);
  eb.enx = ex1 - eb.ex0 + 1;
  eb.eny = ey1 - eb.ey0 + 1;
  return eb;
}

void run_elements(const Dmda &da, int as_written, int rhs_kind, int want, ElemArrays &ea) {
  Ctx *c = da.ctx;
  ea.eb = element_box(da);
  const int64_t nel = (int64_t)ea.eb.enx * ea.eb.eny;
  B2_REQUIRE(nel > 0, "assembly: grid needs at least 2x2 nodes");
  if (want & WANT_K) ea.Ke.alloc((size_t)nel * 64);
  if (want & WANT_F) ea.Fe.alloc((size_t)nel * 8);
  if (want & WANT_KKT) { ea.Ge.alloc((size_t)nel * 32); ea.Ce.alloc((size_t)nel * 16); ea.Qe.alloc((size_t)nel * 16); }
  LaunchScope ls(c, "assembly");
  k_elements<<<(unsigned)((nel + 127) / 128), 128, 0, c->stream>>>(da.M, da.N, ea.eb, as_written, rhs_kind, want, ea.Ke.p, ea.Fe.p, ea.Ge.p, ea.Ce.p, ea.Qe.p);
  check_launch("k_elements");
}

std::shared_ptr<Csr> build_box_matrix(const Dmda &da, const ElemArrays &ea, int dofr, int dofc, int transposed, const double *E) {
  Ctx *c = da.ctx;
  if (da.halo) { // row-partitioned: diagonal + off-diagonal blocks, ghost columns in MPIAIJ order
    BoxEnum e{da.M, da.N, RowBox{da.xs, da.ys, da.xm, da.ym}, ea.eb, dofr, dofc, transposed, E};
    auto A = build_dist(c, e, da.xm * da.ym * dofr, *da.halo, dofc, "spmv");
    A->halo = da.halo;
    A->halo_dof = dofc;
    const int64_t g0 = da.layout->rstart[(size_t)c->rank];
    A->row_gstart = g0 * dofr;
    A->col_gstart = g0 * dofc;
    A->grid_M = da.M; A->grid_N = da.N; A->dof_r = dofr; A->dof_c = dofc;
    A->layout = da.layout;
    return A;
  }
  GridBox g{da.M, da.N, da.xs, da.ys, da.xm, da.ym};
  const int nrows = da.xm * da.ym * dofr, ncols = da.xm * da.ym * dofc;
  DevBuf<int> len((size_t)nrows + 1), rp((size_t)nrows + 1);
  {
    LaunchScope ls(c, "assembly");
    k_box_rowlen<<<grid_for(c, nrows), 256, 0, c->stream>>>(g, dofr, dofc, len.p);
    check_launch("k_box_rowlen");
  }
  int total = 0;
  exclusive_scan_i32(c, len.p, rp.p, nrows, &total);
  auto A = csr_alloc_public(c, nrows, ncols, total);
  B2_CUDA(cudaMemcpyAsync(A->rowptr.p, rp.p, sizeof(int) * ((size_t)nrows + 1), cudaMemcpyDeviceToDevice, c->stream));
  {
    LaunchScope ls(c, "assembly");
    k_box_fill<<<grid_for(c, nrows), 256, 0, c->stream>>>(g, ea.eb, dofr, dofc, transposed, E, A->rowptr.p, A->col.p, A->val.p);
    check_launch("k_box_fill");
  }
  c->sync();
  A->grid_M = da.M; A->grid_N = da.N; A->dof_r = dofr; A->dof_c = dofc;
  A->plan();
  return A;
}

} // namespace

std::shared_ptr<Csr> assemble_stress(const Dmda &da, int as_written, int coeff_kind) {
  ElemArrays ea;
  run_elements(da, as_written, 0, WANT_K | (coeff_kind ? WANT_COEFF : 0), ea);
  auto A = build_box_matrix(da, ea, 2, 2, 0, ea.Ke.p);
  A->tag = "spmv:A";
  csr_try_block_index(*A, 2, 2);
  return A;
}

void assemble_rhs(const Dmda &da, int as_written, int kind, double *f) {
  Ctx *c = da.ctx;
  ElemArrays ea;
  run_elements(da, as_written, kind, WANT_F, ea);
  GridBox g{da.M, da.N, da.xs, da.ys, da.xm, da.ym};
  {
    LaunchScope ls(c, "assembly");
    k_rhs_gather<<<grid_for(c, (int64_t)da.xm * da.ym * 2), 256, 0, c->stream>>>(g, ea.eb, ea.Fe.p, f);
    check_launch("k_rhs_gather");
  }
  c->sync();
}

void assemble_kkt(const Dmda &da, std::shared_ptr<Csr> *Bt, std::shared_ptr<Csr> *B, std::shared_ptr<Csr> *C, std::shared_ptr<Csr> *Q) {
  ElemArrays ea;
  run_elements(da, 0, 0, WANT_KKT, ea);
  if (Bt) { *Bt = build_box_matrix(da, ea, 2, 1, 0, ea.Ge.p); (*Bt)->tag = "spmv:Bt"; csr_try_block_index(**Bt, 2, 1); }
  if (B) {
    *B = build_box_matrix(da, ea, 1, 2, 1, ea.Ge.p); (*B)->tag = "spmv:B";
    // 1x2 blocks: no gain for the plain value stream (0.2206 vs 0.2166 ms at 16M), but the value-dictionary kernel
    // loads both x entries of a block with one 16-byte load
    static const bool bblk = !(getenv("B200SP_B_BLOCK_INDEX") && atoi(getenv("B200SP_B_BLOCK_INDEX")) == 0);
    if (bblk) csr_try_block_index(**B, 1, 2);
  }
  if (C) { *C = build_box_matrix(da, ea, 1, 1, 0, ea.Ce.p); (*C)->tag = "spmv:C"; }
  if (Q) { *Q = build_box_matrix(da, ea, 1, 1, 0, ea.Qe.p); (*Q)->tag = "spmv:Q"; }
}

void assemble_constraints(const Dmda &da, std::shared_ptr<Csr> *B, std::shared_ptr<Csr> *Bt) {
  Ctx *c = da.ctx;
  if (da.halo) throw Error(B200SP_ERR_UNSUPPORTED, "assemble_constraints: the 4 dense constraint rows are assembled on one rank only");
  const int nn = da.M * da.N, n = 2 * nn;
  const int64_t nnz = 2LL * nn + 2LL * n;
  auto Bm = csr_alloc_public(c, 4, n, nnz);
  const int rp[5] = {0, nn, 2 * nn, 2 * nn + n, 2 * nn + 2 * n};
  B2_CUDA(cudaMemcpyAsync(Bm->rowptr.p, rp, sizeof(rp), cudaMemcpyHostToDevice, c->stream));
  {
    LaunchScope ls(c, "assembly");
    k_constraint_rows<<<std::max(1, std::min((nn + 127) / 128, c->num_sms * 16)), 128, 0, c->stream>>>(da.M, da.N, Bm->col.p, Bm->val.p);
    check_launch("k_constraint_rows");
  }
  c->sync();
  Bm->tag = "spmv:Bcon";
  Bm->plan();
  if (Bt) { *Bt = csr_transpose(*Bm); (*Bt)->tag = "spmv:Bcon_t"; }
  if (B) *B = Bm;
}

std::shared_ptr<Csr> interp_q1(Ctx *c, int Mc, int Nc, int dof, int bc) {
  B2_REQUIRE(Mc >= 2 && Nc >= 2 && dof >= 1, "interp_q1: bad grid");
  const int Mf = 2 * Mc - 1, Nf = 2 * Nc - 1;
  const int nrows = Mf * Nf * dof, ncols = Mc * Nc * dof;
  DevBuf<int> len((size_t)nrows + 1), rp((size_t)nrows + 1);
  {
    LaunchScope ls(c, "assembly");
    k_interp_len<<<grid_for(c, nrows), 256, 0, c->stream>>>(Mc, Nc, dof, len.p);
    check_launch("k_interp_len");
  }
  int total = 0;
  exclusive_scan_i32(c, len.p, rp.p, nrows, &total);
  auto P = csr_alloc_public(c, nrows, ncols, total);
  B2_CUDA(cudaMemcpyAsync(P->rowptr.p, rp.p, sizeof(int) * ((size_t)nrows + 1), cudaMemcpyDeviceToDevice, c->stream));
  {
    LaunchScope ls(c, "assembly");
    k_interp_fill<<<grid_for(c, nrows), 256, 0, c->stream>>>(Mc, Nc, dof, bc, P->rowptr.p, P->col.p, P->val.p);
    check_launch("k_interp_fill");
  }
  c->sync();
  P->tag = "spmv:P";
  P->plan();
  return P;
}

std::shared_ptr<Csr> restrict_q1(Ctx *c, int Mc, int Nc, int dof, int bc) {
  B2_REQUIRE(Mc >= 2 && Nc >= 2 && dof >= 1, "restrict_q1: bad grid");
  const int Mf = 2 * Mc - 1, Nf = 2 * Nc - 1;
  const int nrows = Mc * Nc * dof, ncols = Mf * Nf * dof;
  DevBuf<int> len((size_t)nrows + 1), rp((size_t)nrows + 1);
  {
    LaunchScope ls(c, "assembly");
    k_restrict_len<<<grid_for(c, nrows), 256, 0, c->stream>>>(Mc, Nc, dof, len.p);
    check_launch("k_restrict_len");
  }
  int total = 0;
  exclusive_scan_i32(c, len.p, rp.p, nrows, &total);
  auto R = csr_alloc_public(c, nrows, ncols, total);
  B2_CUDA(cudaMemcpyAsync(R->rowptr.p, rp.p, sizeof(int) * ((size_t)nrows + 1), cudaMemcpyDeviceToDevice, c->stream));
  {
    LaunchScope ls(c, "assembly");
    k_restrict_fill<<<grid_for(c, nrows), 256, 0, c->stream>>>(Mc, Nc, dof, bc, R->rowptr.p, R->col.p, R->val.p);
    check_launch("k_restrict_fill");
  }
  c->sync();
  R->tag = "spmv:R";
  R->plan();
  return R;
}

// distributed P (rows: owned fine nodes of `fine`; columns: coarse nodes of `coarse`) and R = P^T
std::shared_ptr<Csr> interp_q1_dist(const Dmda &fine, const Dmda &coarse, int dof, int bc) {
  Ctx *c = fine.ctx;
  B2_REQUIRE(fine.halo && coarse.halo && fine.M == 2 * coarse.M - 1 && fine.N == 2 * coarse.N - 1, "interp_q1_dist: grids do not nest");
  InterpEnum e{coarse.M, coarse.N, dof, bc, RowBox{fine.xs, fine.ys, fine.xm, fine.ym}};
  auto P = build_dist(c, e, fine.xm * fine.ym * dof, *coarse.halo, dof, "spmv:P");
  P->halo = coarse.halo;
  P->halo_dof = dof;
  P->row_gstart = (int64_t)fine.layout->rstart[(size_t)c->rank] * dof;
  P->col_gstart = (int64_t)coarse.layout->rstart[(size_t)c->rank] * dof;
  return P;
}
std::shared_ptr<Csr> restrict_q1_dist(const Dmda &fine, const Dmda &coarse, int dof, int bc) {
  Ctx *c = fine.ctx;
  B2_REQUIRE(fine.halo && coarse.halo && fine.M == 2 * coarse.M - 1 && fine.N == 2 * coarse.N - 1, "restrict_q1_dist: grids do not nest");
  RestrictEnum e{coarse.M, coarse.N, dof, bc, RowBox{coarse.xs, coarse.ys, coarse.xm, coarse.ym}};
  auto R = build_dist(c, e, coarse.xm * coarse.ym * dof, *fine.halo, dof, "spmv:R");
  R->halo = fine.halo;
  R->halo_dof = dof;
  R->row_gstart = (int64_t)coarse.layout->rstart[(size_t)c->rank] * dof;
  R->col_gstart = (int64_t)fine.layout->rstart[(size_t)c->rank] * dof;
  return R;
}

std::vector<int> dmda_bc_ids(const Dmda &da, int dof) {
  // ApplyBC_Laplace id list (:246-262) through the partition map (local row ids), fixing the reference's
  // single-rank-only formula (j*xm+i)*dof+d
  std::vector<int> ids;
  for (int j = da.ys; j < da.ys + da.ym; ++j)
    for (int i = da.xs; i < da.xs + da.xm; ++i)
      if (i == 0 || i == da.M - 1 || j == 0 || j == da.N - 1)
        for (int d = 0; d < dof; ++d) ids.push_back(((j - da.ys) * da.xm + (i - da.xs)) * dof + d);
  return ids;
}

} // namespace b200sp
