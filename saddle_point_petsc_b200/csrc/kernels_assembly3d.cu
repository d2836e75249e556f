// kernels_assembly3d.cu -- device assembly of the 3-D (Q1 hexahedron) KKT discretisation: BASELINE config 4.
//
// The reference is 2-D only (#define DIM 2, include/Discretization.h:8); the 3-D problem is the trilinear analogue of its
// routines, defined in oracle/sp_oracle3d.c (node order, 2x2x2 Gauss rule with the reference's truncated abscissa, 6 x 24
// symmetric-gradient matrix with D~ = diag(2,2,2,1,1,1) w detJ, transposed consumption of Ke, ex42-style gradient /
// stabilisation / mass blocks, all-round Dirichlet).  Same two phases as the 2-D assembly (kernels_assembly.cu):
//   1. element kernels, operations in the oracle's exact order (-fmad=false), results entry-major (SoA over elements);
//   2. CSR stage: one thread per matrix row writes the DMCreateMatrix 27-point box pattern (explicit zeros kept) and sums,
//      for every entry, its <= 8 element contributions in element order (k outer, j, i inner) from +0.0.
// Columns are numbered through a lookup table over the owned box extended by one node layer ("ext box"): owned nodes map
// to their local index, ghost nodes to n_owned + position in the sorted ghost list -- the same kernels serve one rank
// (the table is the natural numbering) and the row-partitioned 3-D DMDA (ghost elements recomputed, no communication).
#include "dev.cuh"
#include "dist.h"
#include <algorithm>

namespace b200sp {

std::shared_ptr<Csr> csr_alloc_public(Ctx *c, int nrows, int ncols, int64_t nnz);

namespace {

struct Grid3 { int M, N, P, xs, ys, zs, xm, ym, zm; };      // global node counts, owned node box
struct ElemBox3 { int ex0, ey0, ez0, enx, eny, enz; };       // element range held in the SoA arrays

__device__ __forceinline__ int n_di(int n) { return (n & 3) >= 2; }
__device__ __forceinline__ int n_dj(int n) { return (n & 3) == 1 || (n & 3) == 2; }
__device__ __forceinline__ int n_dk(int n) { return n >> 2; }
__device__ __forceinline__ double sgn3(int d) { return d ? 1.0 : -1.0; }
// local node number inside its element from the offsets (di, dj, dk): planar DMDAGetElementEqnums order, bottom layer first
__device__ __forceinline__ int local_node3(int di, int dj, int dk) { return (di == 0 ? (dj == 0 ? 0 : 1) : (dj == 0 ? 3 : 2)) + 4 * dk; }

__device__ __forceinline__ void gauss3(int p, double xi[3]) {
  const double g = 0.57735026919;
  const int q = p & 3;
  xi[0] = (q < 2) ? -g : g;
  xi[1] = (q == 0 || q == 3) ? -g : g;
  xi[2] = (p < 4) ? -g : g;
}
__device__ __forceinline__ void q1_3d_Ni(const double xi[3], double Ni[8]) {
#pragma unroll
  for (int n = 0; n < 8; ++n) Ni[n] = 0.125 * (1.0 + sgn3(n_di(n)) * xi[0]) * (1.0 + sgn3(n_dj(n)) * xi[1]) * (1.0 + sgn3(n_dk(n)) * xi[2]);
}
__device__ __forceinline__ void q1_3d_GNi(const double xi[3], double GNi[3][8]) {
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const double sx = sgn3(n_di(n)), sy = sgn3(n_dj(n)), sz = sgn3(n_dk(n));
    const double a = 1.0 + sx * xi[0], b = 1.0 + sy * xi[1], c = 1.0 + sz * xi[2];
    GNi[0][n] = 0.125 * sx * b * c;
    GNi[1][n] = 0.125 * sy * a * c;
    GNi[2][n] = 0.125 * sz * a * b;
  }
}
__device__ __forceinline__ void q1_3d_GNx(const double GNi[3][8], const double *ec, double GNx[3][8], double *detJ) {
  double J[3][3], iJ[3][3];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      double s = 0.0;
#pragma unroll
      for (int n = 0; n < 8; ++n) s += GNi[c][n] * ec[3 * n + d];
      J[c][d] = s;
    }
  const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
  const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
  const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
  const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
  iJ[0][0] = c00 / det;
  iJ[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det;
  iJ[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
  iJ[1][0] = c01 / det;
  iJ[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det;
  iJ[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
  iJ[2][0] = c02 / det;
  iJ[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det;
  iJ[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
#pragma unroll
  for (int n = 0; n < 8; ++n)
#pragma unroll
    for (int c = 0; c < 3; ++c) GNx[c][n] = iJ[c][0] * GNi[0][n] + iJ[c][1] * GNi[1][n] + iJ[c][2] * GNi[2][n];
  *detJ = det;
}
__device__ __forceinline__ void element_coords3(const Grid3 &g, int ei, int ej, int ek, double ec[24]) {
  const double hx = (1.0 - 0.0) / (double)(g.M - 1), hy = (1.0 - 0.0) / (double)(g.N - 1), hz = (1.0 - 0.0) / (double)(g.P - 1);
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    ec[3 * n + 0] = 0.0 + hx * (double)(ei + n_di(n));
    ec[3 * n + 1] = 0.0 + hy * (double)(ej + n_dj(n));
    ec[3 * n + 2] = 0.0 + hz * (double)(ek + n_dk(n));
  }
}
// entry (k, 3a+c) of the 6 x 24 symmetric-gradient matrix from the physical gradient (g0,g1,g2) of node a
__device__ __forceinline__ double strain_entry(int k, int c, double g0, double g1, double g2) {
  switch (k) {
  case 0: return c == 0 ? g0 : 0.0;
  case 1: return c == 1 ? g1 : 0.0;
  case 2: return c == 2 ? g2 : 0.0;
  case 3: return c == 0 ? g1 : (c == 1 ? g0 : 0.0);
  case 4: return c == 0 ? g2 : (c == 2 ? g0 : 0.0);
  default: return c == 1 ? g2 : (c == 2 ? g1 : 0.0);
  }
}
__device__ __forceinline__ void elem_ijk(const ElemBox3 &eb, int64_t e, int *ei, int *ej, int *ek) {
  *ei = eb.ex0 + (int)(e % eb.enx);
  *ej = eb.ey0 + (int)((e / eb.enx) % eb.eny);
  *ek = eb.ez0 + (int)(e / ((int64_t)eb.enx * eb.eny));
}

// stress block: thread (element e, column j) accumulates Ke[i + 24 j], i < 24, over the 8 Gauss points (all six strain rows,
// zeros included, like the oracle's triple loop) and stores it in the row-major position MatSetValuesStencil reads: Ae[j*24 + i]
__global__ void __launch_bounds__(128) k3_elem_K(Grid3 g, ElemBox3 eb, double *__restrict__ Ke) {
  const int64_t nel = (int64_t)eb.enx * eb.eny * eb.enz;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nel * 24) return;
  const int64_t e = t % nel;
  const int j = (int)(t / nel), aj = j / 3, cj = j % 3;
  int ei, ej, ek;
  elem_ijk(eb, e, &ei, &ej, &ek);
  double ec[24];
  element_coords3(g, ei, ej, ek, ec);
  double K[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) K[i] = 0.0;
  for (int p = 0; p < 8; ++p) {
    double xi[3], GNi[3][8], GNx[3][8], detJ, tD[6], Bj[6];
    gauss3(p, xi);
    q1_3d_GNi(xi, GNi);
    q1_3d_GNx(GNi, ec, GNx, &detJ);
    const double coeff = 1.0, w = 1.0;
    tD[0] = tD[1] = tD[2] = 2.0 * w * detJ * coeff;
    tD[3] = tD[4] = tD[5] = w * detJ * coeff;
    double gj0 = 0.0, gj1 = 0.0, gj2 = 0.0;
#pragma unroll
    for (int a = 0; a < 8; ++a)
      if (a == aj) { gj0 = GNx[0][a]; gj1 = GNx[1][a]; gj2 = GNx[2][a]; }
#pragma unroll
    for (int k = 0; k < 6; ++k) Bj[k] = strain_entry(k, cj, gj0, gj1, gj2);
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int k = 0; k < 6; ++k) K[3 * a + c] += strain_entry(k, c, GNx[0][a], GNx[1][a], GNx[2][a]) * tD[k] * Bj[k];
  }
#pragma unroll
  for (int i = 0; i < 24; ++i) Ke[(size_t)(j * 24 + i) * nel + e] = K[i];
}
// gradient block: thread (element, row r = 3a+d): Ge[r*8 + b] -= fac GNx[d][a] N_b
__global__ void __launch_bounds__(128) k3_elem_G(Grid3 g, ElemBox3 eb, double *__restrict__ Ge) {
  const int64_t nel = (int64_t)eb.enx * eb.eny * eb.enz;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nel * 24) return;
  const int64_t e = t % nel;
  const int r = (int)(t / nel), ar = r / 3, dr = r % 3;
  int ei, ej, ek;
  elem_ijk(eb, e, &ei, &ej, &ek);
  double ec[24];
  element_coords3(g, ei, ej, ek, ec);
  double G[8];
#pragma unroll
  for (int b = 0; b < 8; ++b) G[b] = 0.0;
  for (int p = 0; p < 8; ++p) {
    double xi[3], Ni[8], GNi[3][8], GNx[3][8], detJ;
    gauss3(p, xi);
    q1_3d_Ni(xi, Ni);
    q1_3d_GNi(xi, GNi);
    q1_3d_GNx(GNi, ec, GNx, &detJ);
    const double fac = 1.0 * detJ;
    double gv = 0.0;
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int d = 0; d < 3; ++d)
        if (a == ar && d == dr) gv = GNx[d][a];
#pragma unroll
    for (int b = 0; b < 8; ++b) G[b] -= fac * gv * Ni[b];
  }
#pragma unroll
  for (int b = 0; b < 8; ++b) Ge[(size_t)(r * 8 + b) * nel + e] = G[b];
}
// stabilisation and (minus) mass blocks: thread (element, row a)
__global__ void __launch_bounds__(128) k3_elem_CQ(Grid3 g, ElemBox3 eb, double *__restrict__ Ce, double *__restrict__ Qe) {
  const int64_t nel = (int64_t)eb.enx * eb.eny * eb.enz;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nel * 8) return;
  const int64_t e = t % nel;
  const int ar = (int)(t / nel);
  int ei, ej, ek;
  elem_ijk(eb, e, &ei, &ej, &ek);
  double ec[24];
  element_coords3(g, ei, ej, ek, ec);
  double Cv[8], Qv[8];
#pragma unroll
  for (int b = 0; b < 8; ++b) { Cv[b] = 0.0; Qv[b] = 0.0; }
  for (int p = 0; p < 8; ++p) {
    double xi[3], Ni[8], GNi[3][8], GNx[3][8], detJ;
    gauss3(p, xi);
    q1_3d_Ni(xi, Ni);
    q1_3d_GNi(xi, GNi);
    q1_3d_GNx(GNi, ec, GNx, &detJ);
    const double fac = 1.0 * detJ;
    double na = 0.0;
#pragma unroll
    for (int a = 0; a < 8; ++a)
      if (a == ar) na = Ni[a];
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      Cv[b] -= fac * (na * Ni[b] - 0.015625);
      Qv[b] -= fac * (na * Ni[b]);
    }
  }
#pragma unroll
  for (int b = 0; b < 8; ++b) { Ce[(size_t)(ar * 8 + b) * nel + e] = Cv[b]; Qe[(size_t)(ar * 8 + b) * nel + e] = Qv[b]; }
}
// right-hand side: thread per element
__global__ void __launch_bounds__(128) k3_elem_F(Grid3 g, ElemBox3 eb, int kind, double *__restrict__ Fe) {
  const int64_t nel = (int64_t)eb.enx * eb.eny * eb.enz;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nel) return;
  int ei, ej, ek;
  elem_ijk(eb, e, &ei, &ej, &ek);
  double ec[24];
  element_coords3(g, ei, ej, ek, ec);
  double F[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) F[i] = 0.0;
  for (int p = 0; p < 8; ++p) {
    double xi[3], Ni[8], GNi[3][8], GNx[3][8], detJ, f_p[3];
    gauss3(p, xi);
    q1_3d_Ni(xi, Ni);
    q1_3d_GNi(xi, GNi);
    q1_3d_GNx(GNi, ec, GNx, &detJ);
    const double fac = 1.0 * detJ;
    if (kind == 0) { f_p[0] = 1.0; f_p[1] = 2.0; f_p[2] = 3.0; }
    else {
      double xp = 0.0, yp = 0.0;
#pragma unroll
      for (int n = 0; n < 8; ++n) { xp += Ni[n] * ec[3 * n]; yp += Ni[n] * ec[3 * n + 1]; }
      f_p[0] = 2.0 * yp - 1.0;
      f_p[1] = 1.0 - 2.0 * xp;
      f_p[2] = 0.0;
    }
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int c = 0; c < 3; ++c) F[3 * a + c] += fac * Ni[a] * f_p[c];
  }
#pragma unroll
  for (int i = 0; i < 24; ++i) Fe[(size_t)i * nel + e] = F[i];
}

__device__ __forceinline__ void owned_ijk(const Grid3 &g, int node, int *i, int *j, int *k) {
  *i = g.xs + node % g.xm;
  *j = g.ys + (node / g.xm) % g.ym;
  *k = g.zs + node / (g.xm * g.ym);
}
__device__ __forceinline__ int clip_lo(int v) { return v > 0 ? v - 1 : 0; }
__device__ __forceinline__ int clip_hi(int v, int n) { return v < n - 1 ? v + 1 : n - 1; }

__global__ void __launch_bounds__(256) k3_box_rowlen(Grid3 g, int dofr, int dofc, int *len) {
  const int nrows = g.xm * g.ym * g.zm * dofr;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x) {
    int i, j, k;
    owned_ijk(g, r / dofr, &i, &j, &k);
    len[r] = dofc * (clip_hi(i, g.M) - clip_lo(i) + 1) * (clip_hi(j, g.N) - clip_lo(j) + 1) * (clip_hi(k, g.P) - clip_lo(k) + 1);
  }
}
// lut: local column node id of every node of the owned box extended by one layer ((xm+2) x (ym+2) x (zm+2), -1 outside the domain)
__global__ void __launch_bounds__(128) k3_box_fill(Grid3 g, ElemBox3 eb, int dofr, int dofc, int transposed, const double *__restrict__ E,
                                                   const int *__restrict__ lut, const int *__restrict__ rowptr, int *col, double *val) {
  const int nrows = g.xm * g.ym * g.zm * dofr;
  const int64_t nel = (int64_t)eb.enx * eb.eny * eb.enz;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x) {
    const int cr = r % dofr;
    int i, j, k;
    owned_ijk(g, r / dofr, &i, &j, &k);
    int p = rowptr[r];
    for (int kk = clip_lo(k); kk <= clip_hi(k, g.P); ++kk)
      for (int jj = clip_lo(j); jj <= clip_hi(j, g.N); ++jj)
        for (int ii = clip_lo(i); ii <= clip_hi(i, g.M); ++ii) {
          const int cnode = lut[((kk - g.zs + 1) * (g.ym + 2) + (jj - g.ys + 1)) * (g.xm + 2) + (ii - g.xs + 1)];
          for (int cc = 0; cc < dofc; ++cc) {
            double acc = 0.0;
            for (int ek = max(max(k, kk) - 1, 0); ek <= min(min(k, kk), g.P - 2); ++ek)
              for (int ej = max(max(j, jj) - 1, 0); ej <= min(min(j, jj), g.N - 2); ++ej)
                for (int ei = max(max(i, ii) - 1, 0); ei <= min(min(i, ii), g.M - 2); ++ei) {
                  const int la = local_node3(i - ei, j - ej, k - ek), lb = local_node3(ii - ei, jj - ej, kk - ek);
                  const int entry = transposed ? (lb * dofc + cc) * (8 * dofr) + (la * dofr + cr) : (la * dofr + cr) * (8 * dofc) + (lb * dofc + cc);
                  const int64_t e = ((int64_t)(ek - eb.ez0) * eb.eny + (ej - eb.ey0)) * eb.enx + (ei - eb.ex0);
                  acc += E[(size_t)entry * nel + e];
                }
            col[p] = cnode * dofc + cc;
            val[p] = acc;
            ++p;
          }
        }
  }
}
__global__ void __launch_bounds__(256) k3_rhs_gather(Grid3 g, ElemBox3 eb, const double *__restrict__ Fe, double *__restrict__ f) {
  const int nrows = g.xm * g.ym * g.zm * 3;
  const int64_t nel = (int64_t)eb.enx * eb.eny * eb.enz;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x) {
    const int c = r % 3;
    int i, j, k;
    owned_ijk(g, r / 3, &i, &j, &k);
    double acc = 0.0;
    for (int ek = max(k - 1, 0); ek <= min(k, g.P - 2); ++ek)
      for (int ej = max(j - 1, 0); ej <= min(j, g.N - 2); ++ej)
        for (int ei = max(i - 1, 0); ei <= min(i, g.M - 2); ++ei) {
          const int entry = local_node3(i - ei, j - ej, k - ek) * 3 + c;
          const int64_t e = ((int64_t)(ek - eb.ez0) * eb.eny + (ej - eb.ey0)) * eb.enx + (ei - eb.ex0);
          acc += Fe[(size_t)entry * nel + e];
        }
    f[r] = acc;
  }
}

inline int grid_for3(Ctx *c, int64_t n, int threads) {
  int64_t gsz = (n + threads - 1) / threads;
  return (int)std::max<int64_t>(1, std::min<int64_t>(gsz, (int64_t)c->num_sms * 32));
}
Grid3 grid_of(const Dmda3 &da) { return Grid3{da.M, da.N, da.P, da.xs, da.ys, da.zs, da.xm, da.ym, da.zm}; }
ElemBox3 element_box3(const Dmda3 &da) { // ghost-element recomputation: every element touching an owned node
  ElemBox3 eb;
  eb.ex0 = std::max(da.xs - 1, 0); eb.ey0 = std::max(da.ys - 1, 0); eb.ez0 = std::max(da.zs - 1, 0);
  eb.enx = std::min(da.xs + da.xm - 1, da.M - 2) - eb.ex0 + 1;
  eb.eny = std::min(da.ys + da.ym - 1, da.N - 2) - eb.ey0 + 1;
  eb.enz = std::min(da.zs + da.zm - 1, da.P - 2) - eb.ez0 + 1;
  return eb;
}

std::shared_ptr<Csr> build_box_matrix3(const Dmda3 &da, const ElemBox3 &eb, int dofr, int dofc, int transposed, const double *E, const char *tag) {
  Ctx *c = da.ctx;
  const Grid3 g = grid_of(da);
  const int nown = da.xm * da.ym * da.zm, nrows = nown * dofr, ncols = nown * dofc;
  DevBuf<int> len((size_t)nrows + 1), rp((size_t)nrows + 1);
  {
    LaunchScope ls(c, "assembly");
    k3_box_rowlen<<<grid_for3(c, nrows, 256), 256, 0, c->stream>>>(g, dofr, dofc, len.p);
    check_launch("k3_box_rowlen");
  }
  // 32-bit row pointers: check the size BEFORE the int32 scan can overflow
  int64_t nnz64 = 0;
  {
    std::vector<int> hl((size_t)nrows);
    B2_CUDA(cudaMemcpyAsync(hl.data(), len.p, sizeof(int) * (size_t)nrows, cudaMemcpyDeviceToHost, c->stream));
    c->sync();
    for (int v : hl) nnz64 += v;
  }
  if (nnz64 >= (int64_t)2147483647 - CSR_PAD)
    throw Error(B200SP_ERR_UNSUPPORTED, "3-D matrix with " + std::to_string(nnz64) + " stored entries on one rank: 32-bit row pointers hold < 2^31; use more ranks");
  int total = 0;
  exclusive_scan_i32(c, len.p, rp.p, nrows, &total);
  auto A = csr_alloc_public(c, nrows, ncols, total);
  B2_CUDA(cudaMemcpyAsync(A->rowptr.p, rp.p, sizeof(int) * ((size_t)nrows + 1), cudaMemcpyDeviceToDevice, c->stream));
  {
    LaunchScope ls(c, "assembly");
    k3_box_fill<<<grid_for3(c, nrows, 128), 128, 0, c->stream>>>(g, eb, dofr, dofc, transposed, E, da.lut.p, A->rowptr.p, A->col.p, A->val.p);
    check_launch("k3_box_fill");
  }
  c->sync();
  A->dof_r = dofr; A->dof_c = dofc;
  A->tag = tag;
  if (da.halo) { A->halo = da.halo; A->halo_dof = dofc; A->row_gstart = da.g0 * dofr; A->col_gstart = da.g0 * dofc; }
  A->plan();
  if (dofr * dofc > 1) csr_try_block_index(*A, dofr, dofc); // 3 x 3 / 3 x 1 / 1 x 3 node blocks -> warp-per-node SpMV
  return A;
}

} // namespace

// the ext-box lookup table of a 3-D DMDA: local column node id of every node in the owned box extended by one layer
void dmda3_build_lut(Dmda3 &da, const std::vector<int> &host_lut) {
  da.lut.alloc(host_lut.size() + 1);
  B2_CUDA(cudaMemcpyAsync(da.lut.p, host_lut.data(), sizeof(int) * host_lut.size(), cudaMemcpyHostToDevice, da.ctx->stream));
  da.ctx->sync();
}

std::shared_ptr<Csr> assemble3_stress(const Dmda3 &da) {
  Ctx *c = da.ctx;
  const ElemBox3 eb = element_box3(da);
  const int64_t nel = (int64_t)eb.enx * eb.eny * eb.enz;
  B2_REQUIRE(nel > 0, "3-D assembly: grid needs at least 2 x 2 x 2 nodes");
  DevBuf<double> Ke((size_t)nel * 576);
  {
    LaunchScope ls(c, "assembly");
    k3_elem_K<<<(unsigned)((nel * 24 + 127) / 128), 128, 0, c->stream>>>(grid_of(da), eb, Ke.p);
    check_launch("k3_elem_K");
  }
  auto A = build_box_matrix3(da, eb, 3, 3, 0, Ke.p, "spmv:A");
  return A;
}
void assemble3_rhs(const Dmda3 &da, int kind, double *f) {
  Ctx *c = da.ctx;
  const ElemBox3 eb = element_box3(da);
  const int64_t nel = (int64_t)eb.enx * eb.eny * eb.enz;
  DevBuf<double> Fe((size_t)nel * 24);
  {
    LaunchScope ls(c, "assembly");
    k3_elem_F<<<(unsigned)((nel + 127) / 128), 128, 0, c->stream>>>(grid_of(da), eb, kind, Fe.p);
    check_launch("k3_elem_F");
    k3_rhs_gather<<<grid_for3(c, (int64_t)da.xm * da.ym * da.zm * 3, 256), 256, 0, c->stream>>>(grid_of(da), eb, Fe.p, f);
    check_launch("k3_rhs_gather");
  }
  c->sync();
}
void assemble3_kkt(const Dmda3 &da, std::shared_ptr<Csr> *Bt, std::shared_ptr<Csr> *B, std::shared_ptr<Csr> *C, std::shared_ptr<Csr> *Q) {
  Ctx *c = da.ctx;
  const ElemBox3 eb = element_box3(da);
  const int64_t nel = (int64_t)eb.enx * eb.eny * eb.enz;
  {
    DevBuf<double> Ge((size_t)nel * 192);
    {
      LaunchScope ls(c, "assembly");
      k3_elem_G<<<(unsigned)((nel * 24 + 127) / 128), 128, 0, c->stream>>>(grid_of(da), eb, Ge.p);
      check_launch("k3_elem_G");
    }
    if (Bt) *Bt = build_box_matrix3(da, eb, 3, 1, 0, Ge.p, "spmv:Bt");
    if (B) *B = build_box_matrix3(da, eb, 1, 3, 1, Ge.p, "spmv:B");
  }
  DevBuf<double> Ce((size_t)nel * 64), Qe((size_t)nel * 64);
  {
    LaunchScope ls(c, "assembly");
    k3_elem_CQ<<<(unsigned)((nel * 8 + 127) / 128), 128, 0, c->stream>>>(grid_of(da), eb, Ce.p, Qe.p);
    check_launch("k3_elem_CQ");
  }
  if (C) *C = build_box_matrix3(da, eb, 1, 1, 0, Ce.p, "spmv:C");
  if (Q) *Q = build_box_matrix3(da, eb, 1, 1, 0, Qe.p, "spmv:Q");
}
// ApplyBC_Laplace in 3-D: local row ids of the owned boundary nodes, all components, ascending
std::vector<int> dmda3_bc_ids(const Dmda3 &da, int dof) {
  std::vector<int> ids;
  for (int k = 0; k < da.zm; ++k)
    for (int j = 0; j < da.ym; ++j)
      for (int i = 0; i < da.xm; ++i) {
        const int gi = da.xs + i, gj = da.ys + j, gk = da.zs + k;
        if (gi == 0 || gi == da.M - 1 || gj == 0 || gj == da.N - 1 || gk == 0 || gk == da.P - 1)
          for (int d = 0; d < dof; ++d) ids.push_back(((k * da.ym + j) * da.xm + i) * dof + d);
      }
  return ids;
}

} // namespace b200sp
