/*
 * sp_oracle3d.c -- CPU ORACLE, 3-D part (test infrastructure, NOT the product).
 *
 * BASELINE config 4 asks for a 3-D Stokes-type KKT system; the reference is 2-D only (#define DIM 2,
 * include/Discretization.h:8), so there is no reference code to follow and nothing to pin against: "parity unpinned".
 * The 3-D discretisation is DEFINED here as the trilinear (Q1 hexahedron) analogue of the reference's 2-D routines,
 * keeping every convention the 2-D code fixes:
 *   nodes      planar order of DMDAGetElementEqnums (src/Discretization.c:377-395): (0,0),(0,1),(1,1),(1,0) in (di,dj),
 *              bottom layer (dk = 0) first, then the top layer
 *   quadrature 2x2x2 Gauss with the reference's truncated abscissa 0.57735026919 (:52-55), weights 1, planar point
 *              order of ConstructGaussQuadratureQ12D (:49-63) for zeta = -g, then zeta = +g
 *   A          FormStressOperatorQ12D (:293-332) with the 6 x 24 symmetric-gradient matrix and D~ = diag(2,2,2,1,1,1) w detJ;
 *              accumulated as Ke[i + 24 j], consumed row-major like MatSetValuesStencil does (:165, :327)
 *   f          FormLaplaceRHSQ12D (:334-374): Fe[3a+c] += fac N_a f_c
 *   B^T, C, Q  the ex42/ex43 blocks of or_element_kkt (sp_oracle.c) with 1/64 as the projection constant
 *   BC         homogeneous Dirichlet on every boundary node, all three components (ApplyBC_Laplace, :229-274)
 *   numbering  natural: node (i,j,k) -> (k N + j) M + i, dof-interleaved; elements looped k outer, j, i inner
 * The CUDA assembly (saddle_point_petsc_b200/csrc/kernels_assembly3d.cu) mirrors these operations one for one.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "sp_oracle.h"

static const double G3 = 0.57735026919;
/* node n of the element: offsets (di,dj,dk) and the signs of its shape function */
static const int DI3[8] = {0, 0, 1, 1, 0, 0, 1, 1}, DJ3[8] = {0, 1, 1, 0, 0, 1, 1, 0}, DK3[8] = {0, 0, 0, 0, 1, 1, 1, 1};

static void gauss3(int p, double xi[3]) {
  const int q = p & 3;
  xi[0] = (q < 2) ? -G3 : G3;
  xi[1] = (q == 0 || q == 3) ? -G3 : G3;
  xi[2] = (p < 4) ? -G3 : G3;
}
static double sgn3(int d) { return d ? 1.0 : -1.0; }
static void q1_3d_Ni(const double xi[3], double Ni[8]) {
  for (int n = 0; n < 8; ++n) Ni[n] = 0.125 * (1.0 + sgn3(DI3[n]) * xi[0]) * (1.0 + sgn3(DJ3[n]) * xi[1]) * (1.0 + sgn3(DK3[n]) * xi[2]);
}
static void q1_3d_GNi(const double xi[3], double GNi[3][8]) {
  for (int n = 0; n < 8; ++n) {
    const double sx = sgn3(DI3[n]), sy = sgn3(DJ3[n]), sz = sgn3(DK3[n]);
    const double a = 1.0 + sx * xi[0], b = 1.0 + sy * xi[1], c = 1.0 + sz * xi[2];
    GNi[0][n] = 0.125 * sx * b * c;
    GNi[1][n] = 0.125 * sy * a * c;
    GNi[2][n] = 0.125 * sz * a * b;
  }
}
/* Jacobian J[c][d] = sum_n GNi[c][n] x_n[d], inverse by cofactors, GNx = invJ * GNi (ConstructQ12D_GNx in 3-D) */
static void q1_3d_GNx(double GNi[3][8], const double *ec, double GNx[3][8], double *detJ) {
  double J[3][3], iJ[3][3];
  for (int c = 0; c < 3; ++c)
    for (int d = 0; d < 3; ++d) {
      double s = 0.0;
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
  for (int n = 0; n < 8; ++n)
    for (int c = 0; c < 3; ++c) GNx[c][n] = iJ[c][0] * GNi[0][n] + iJ[c][1] * GNi[1][n] + iJ[c][2] * GNi[2][n];
  *detJ = det;
}
/* DMDASetUniformCoordinates(0,1)^3 + the (intended) GetElementCoords */
void or3_element_coords(int M, int N, int P, int ei, int ej, int ek, double ec[24]) {
  const double hx = (1.0 - 0.0) / (double)(M - 1), hy = (1.0 - 0.0) / (double)(N - 1), hz = (1.0 - 0.0) / (double)(P - 1);
  for (int n = 0; n < 8; ++n) {
    ec[3 * n + 0] = 0.0 + hx * (double)(ei + DI3[n]);
    ec[3 * n + 1] = 0.0 + hy * (double)(ej + DJ3[n]);
    ec[3 * n + 2] = 0.0 + hz * (double)(ek + DK3[n]);
  }
}
/* the 6 x 24 symmetric-gradient matrix of one Gauss point: rows exx, eyy, ezz, gxy, gxz, gyz; column 3a+c */
static void strain_matrix(double GNx[3][8], double B[6][24]) {
  memset(B, 0, sizeof(double) * 6 * 24);
  for (int a = 0; a < 8; ++a) {
    B[0][3 * a + 0] = GNx[0][a];
    B[1][3 * a + 1] = GNx[1][a];
    B[2][3 * a + 2] = GNx[2][a];
    B[3][3 * a + 0] = GNx[1][a]; B[3][3 * a + 1] = GNx[0][a];
    B[4][3 * a + 0] = GNx[2][a]; B[4][3 * a + 2] = GNx[0][a];
    B[5][3 * a + 1] = GNx[2][a]; B[5][3 * a + 2] = GNx[1][a];
  }
}
void or3_element_stress(const double ec[24], double Ke[576]) { /* += */
  for (int p = 0; p < 8; ++p) {
    double xi[3], GNi[3][8], GNx[3][8], detJ, B[6][24], tD[6];
    gauss3(p, xi);
    q1_3d_GNi(xi, GNi);
    q1_3d_GNx(GNi, ec, GNx, &detJ);
    strain_matrix(GNx, B);
    const double coeff = 1.0, w = 1.0;
    tD[0] = tD[1] = tD[2] = 2.0 * w * detJ * coeff;
    tD[3] = tD[4] = tD[5] = w * detJ * coeff;
    for (int i = 0; i < 24; ++i)
      for (int j = 0; j < 24; ++j)
        for (int k = 0; k < 6; ++k) Ke[i + 24 * j] += B[k][i] * tD[k] * B[k][j];
  }
}
/* kind 0: constant body force (1,2,3); kind 1: rotational force about the z axis through the centre (2y-1, 1-2x, 0)
 * evaluated at the physical Gauss point (the 3-D analogue of the 2-D benchmark force) */
void or3_element_rhs(const double ec[24], int kind, double Fe[24]) { /* += */
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
      for (int n = 0; n < 8; ++n) { xp += Ni[n] * ec[3 * n]; yp += Ni[n] * ec[3 * n + 1]; }
      f_p[0] = 2.0 * yp - 1.0;
      f_p[1] = 1.0 - 2.0 * xp;
      f_p[2] = 0.0;
    }
    for (int a = 0; a < 8; ++a)
      for (int c = 0; c < 3; ++c) Fe[3 * a + c] += fac * Ni[a] * f_p[c];
  }
}
void or3_element_kkt(const double ec[24], double Ge[192], double Ce[64], double Qe[64]) { /* += */
  for (int p = 0; p < 8; ++p) {
    double xi[3], Ni[8], GNi[3][8], GNx[3][8], detJ;
    gauss3(p, xi);
    q1_3d_Ni(xi, Ni);
    q1_3d_GNi(xi, GNi);
    q1_3d_GNx(GNi, ec, GNx, &detJ);
    const double fac = 1.0 * detJ;
    for (int a = 0; a < 8; ++a)
      for (int d = 0; d < 3; ++d)
        for (int b = 0; b < 8; ++b) Ge[(3 * a + d) * 8 + b] -= fac * GNx[d][a] * Ni[b];
    for (int a = 0; a < 8; ++a)
      for (int b = 0; b < 8; ++b) {
        Ce[a * 8 + b] -= fac * (Ni[a] * Ni[b] - 0.015625);
        Qe[a * 8 + b] -= fac * (Ni[a] * Ni[b]);
      }
  }
}

/* DMCreateMatrix on a 3-D box-stencil DMDA: all nodes of the clipped 3x3x3 box, ascending natural index, explicit zeros */
static OrCsr *box_pattern3(int M, int N, int P, int dofr, int dofc) {
  long nnz = (long)dofr * dofc * (3L * M - 2) * (3L * N - 2) * (3L * P - 2);
  if (nnz >= 2147483647L) { fprintf(stderr, "sp_oracle3d: matrix too large for 32-bit row pointers\n"); abort(); }
  OrCsr *A = or_csr_alloc(dofr * M * N * P, dofc * M * N * P, nnz);
  long p = 0;
  for (int k = 0; k < P; ++k)
    for (int j = 0; j < N; ++j)
      for (int i = 0; i < M; ++i)
        for (int c = 0; c < dofr; ++c) {
          for (int kk = (k > 0 ? k - 1 : 0); kk <= (k < P - 1 ? k + 1 : P - 1); ++kk)
            for (int jj = (j > 0 ? j - 1 : 0); jj <= (j < N - 1 ? j + 1 : N - 1); ++jj)
              for (int ii = (i > 0 ? i - 1 : 0); ii <= (i < M - 1 ? i + 1 : M - 1); ++ii)
                for (int cc = 0; cc < dofc; ++cc) A->col[p++] = ((kk * N + jj) * M + ii) * dofc + cc;
          A->rowptr[((k * N + j) * M + i) * dofr + c + 1] = (int)p;
        }
  for (long t = 0; t < nnz; ++t) A->val[t] = 0.0;
  return A;
}
static void add_value3(OrCsr *A, int row, int col, double v) {
  int lo = A->rowptr[row], hi = A->rowptr[row + 1];
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (A->col[mid] < col) lo = mid + 1; else hi = mid;
  }
  if (lo >= A->rowptr[row + 1] || A->col[lo] != col) { fprintf(stderr, "sp_oracle3d: entry (%d,%d) outside preallocation\n", row, col); abort(); }
  A->val[lo] += v;
}
static void element_nodes3(int M, int N, int ei, int ej, int ek, int nd[8]) {
  for (int n = 0; n < 8; ++n) nd[n] = ((ek + DK3[n]) * N + ej + DJ3[n]) * M + ei + DI3[n];
}

OrCsr *or3_assemble_A(int M, int N, int P) {
  OrCsr *A = box_pattern3(M, N, P, 3, 3);
  for (int ek = 0; ek < P - 1; ++ek)
    for (int ej = 0; ej < N - 1; ++ej)
      for (int ei = 0; ei < M - 1; ++ei) {
        double ec[24], Ae[576];
        int nd[8];
        or3_element_coords(M, N, P, ei, ej, ek, ec);
        memset(Ae, 0, sizeof(Ae));
        or3_element_stress(ec, Ae);
        element_nodes3(M, N, ei, ej, ek, nd);
        for (int a = 0; a < 24; ++a)
          for (int b = 0; b < 24; ++b) add_value3(A, nd[a / 3] * 3 + a % 3, nd[b / 3] * 3 + b % 3, Ae[a * 24 + b]); /* row-major read */
      }
  return A;
}
void or3_assemble_rhs(int M, int N, int P, int kind, double *f) {
  memset(f, 0, sizeof(double) * 3 * (size_t)M * N * P);
  for (int ek = 0; ek < P - 1; ++ek)
    for (int ej = 0; ej < N - 1; ++ej)
      for (int ei = 0; ei < M - 1; ++ei) {
        double ec[24], Fe[24];
        int nd[8];
        or3_element_coords(M, N, P, ei, ej, ek, ec);
        memset(Fe, 0, sizeof(Fe));
        or3_element_rhs(ec, kind, Fe);
        element_nodes3(M, N, ei, ej, ek, nd);
        for (int a = 0; a < 8; ++a)
          for (int c = 0; c < 3; ++c) f[nd[a] * 3 + c] += Fe[3 * a + c];
      }
}
void or3_assemble_kkt(int M, int N, int P, OrCsr **pBt, OrCsr **pB, OrCsr **pC, OrCsr **pQ) {
  OrCsr *Bt = box_pattern3(M, N, P, 3, 1), *B = box_pattern3(M, N, P, 1, 3), *C = box_pattern3(M, N, P, 1, 1), *Q = box_pattern3(M, N, P, 1, 1);
  for (int ek = 0; ek < P - 1; ++ek)
    for (int ej = 0; ej < N - 1; ++ej)
      for (int ei = 0; ei < M - 1; ++ei) {
        double ec[24], Ge[192], Ce[64], Qe[64];
        int nd[8];
        or3_element_coords(M, N, P, ei, ej, ek, ec);
        memset(Ge, 0, sizeof(Ge)); memset(Ce, 0, sizeof(Ce)); memset(Qe, 0, sizeof(Qe));
        or3_element_kkt(ec, Ge, Ce, Qe);
        element_nodes3(M, N, ei, ej, ek, nd);
        for (int a = 0; a < 24; ++a)
          for (int b = 0; b < 8; ++b) {
            add_value3(Bt, nd[a / 3] * 3 + a % 3, nd[b], Ge[a * 8 + b]);
            add_value3(B, nd[b], nd[a / 3] * 3 + a % 3, Ge[a * 8 + b]);
          }
        for (int a = 0; a < 8; ++a)
          for (int b = 0; b < 8; ++b) {
            add_value3(C, nd[a], nd[b], Ce[a * 8 + b]);
            add_value3(Q, nd[a], nd[b], Qe[a * 8 + b]);
          }
      }
  *pBt = Bt; *pB = B; *pC = C; *pQ = Q;
}
int or3_bc_ids(int M, int N, int P, int dof, int *ids) {
  int n = 0;
  for (int k = 0; k < P; ++k)
    for (int j = 0; j < N; ++j)
      for (int i = 0; i < M; ++i)
        if (i == 0 || i == M - 1 || j == 0 || j == N - 1 || k == 0 || k == P - 1)
          for (int d = 0; d < dof; ++d) { if (ids) ids[n] = ((k * N + j) * M + i) * dof + d; ++n; }
  return n;
}

/* DMDACreate3d(PETSC_DECIDE x 3): process grid (PETSc da3.c), ownership M/m + (M % m > i) per direction, rank r at
 * (r % m, (r / m) % n, r / (m n)); PETSc global numbering = rank-contiguous, x fastest inside a rank. */
void or3_dmda_proc_grid(int M, int N, int P, int size, int *pm, int *pn, int *pp) {
  int n = (int)(0.5 + pow(((double)N * N) * ((double)size) / ((double)P * M), 1.0 / 3.0)), m, p = 1;
  if (!n) n = 1;
  while (n > 0) { int pmn = size / n; if (n * pmn == size) break; n--; }
  if (!n) n = 1;
  m = (int)(0.5 + sqrt(((double)M) * ((double)size) / ((double)P * n)));
  if (!m) m = 1;
  while (m > 0) { p = size / (m * n); if (m * n * p == size) break; m--; }
  if (M > P && m < p) { int t = m; m = p; p = t; }
  *pm = m; *pn = n; *pp = p;
}
void or3_dmda_natural_to_petsc(int M, int N, int P, int size, int *node_map, int *node_owner) {
  int m, n, p;
  or3_dmda_proc_grid(M, N, P, size, &m, &n, &p);
  int *lx = (int *)malloc(sizeof(int) * (size_t)m), *ly = (int *)malloc(sizeof(int) * (size_t)n), *lz = (int *)malloc(sizeof(int) * (size_t)p);
  or_dmda_ownership(M, m, lx);
  or_dmda_ownership(N, n, ly);
  or_dmda_ownership(P, p, lz);
  int start = 0;
  for (int r = 0; r < size; ++r) {
    int pi = r % m, pj = (r / m) % n, pk = r / (m * n), xs = 0, ys = 0, zs = 0;
    for (int i = 0; i < pi; ++i) xs += lx[i];
    for (int j = 0; j < pj; ++j) ys += ly[j];
    for (int k = 0; k < pk; ++k) zs += lz[k];
    for (int k = 0; k < lz[pk]; ++k)
      for (int j = 0; j < ly[pj]; ++j)
        for (int i = 0; i < lx[pi]; ++i) {
          int nat = ((zs + k) * N + ys + j) * M + xs + i;
          node_map[nat] = start + (k * ly[pj] + j) * lx[pi] + i;
          if (node_owner) node_owner[nat] = r;
        }
    start += lx[pi] * ly[pj] * lz[pk];
  }
  free(lx); free(ly); free(lz);
}
