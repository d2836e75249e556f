/*
 * sp_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).  See sp_oracle.h.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -fPIC -shared (oracle/Makefile).
 * -ffp-contract=off is REQUIRED: every a*b+c below is a separately rounded multiply and add
 * unless written as fma(); the CUDA kernels mirror the same choice operation by operation.
 *
 * Citations "Discretization.c:NNN" refer to /root/reference/src/Discretization.c.
 */
#include "sp_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MINRES_HAPTOL 1e-18 /* KSPMINRES haptol: a tiny negative r.z is rounding, not an indefinite PC */
#define CHUNK 4096 /* fixed reduction chunk: dot/norm results do not depend on the thread count */

static void *xmalloc(size_t n) {
  void *p = malloc(n ? n : 1);
  if (!p) { fprintf(stderr, "sp_oracle: out of memory (%zu bytes)\n", n); abort(); }
  return p;
}
static void *xcalloc(size_t n, size_t s) {
  void *p = calloc(n ? n : 1, s);
  if (!p) { fprintf(stderr, "sp_oracle: out of memory\n"); abort(); }
  return p;
}

void or_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
int or_get_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ============================================================ vectors ===== */
double or_dot(int n, const double *x, const double *y) {
  int nch = (n + CHUNK - 1) / CHUNK;
  double *part = (double *)xmalloc(sizeof(double) * (size_t)nch);
#pragma omp parallel for schedule(static)
  for (int c = 0; c < nch; ++c) {
    int lo = c * CHUNK, hi = lo + CHUNK > n ? n : lo + CHUNK;
    double s = 0.0;
    for (int i = lo; i < hi; ++i) s += x[i] * y[i];
    part[c] = s;
  }
  double s = 0.0;
  for (int c = 0; c < nch; ++c) s += part[c];
  free(part);
  return s;
}
double or_norm2(int n, const double *x) { return sqrt(or_dot(n, x, x)); }
static void v_copy(int n, const double *x, double *y) { memcpy(y, x, sizeof(double) * (size_t)n); }
static void v_zero(int n, double *y) { memset(y, 0, sizeof(double) * (size_t)n); }
static void v_axpy(int n, double a, const double *x, double *y) { /* y += a x */
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) y[i] += a * x[i];
}
static void v_aypx(int n, double a, const double *x, double *y) { /* y = x + a y */
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) y[i] = x[i] + a * y[i];
}
static void v_scale(int n, double a, double *y) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) y[i] *= a;
}
static int v_bad(double r) { return isnan(r) || isinf(r); }

void or_hash_vector(int n, double *v) {
  for (int i = 0; i < n; ++i) {
    unsigned int h = (unsigned int)i * 2654435761u;
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13;
    v[i] = 0.5 + (double)(h >> 8) / 16777216.0; /* exact: 24-bit integer / 2^24 */
  }
}

/* ================================================================ CSR ===== */
OrCsr *or_csr_alloc(int nrows, int ncols, long nnz) {
  OrCsr *A = (OrCsr *)xmalloc(sizeof(OrCsr));
  A->nrows = nrows; A->ncols = ncols;
  A->rowptr = (int *)xcalloc((size_t)nrows + 1, sizeof(int));
  A->col = (int *)xmalloc(sizeof(int) * (size_t)nnz);
  A->val = (double *)xcalloc((size_t)nnz, sizeof(double));
  return A;
}
void or_csr_free(OrCsr *A) {
  if (!A) return;
  free(A->rowptr); free(A->col); free(A->val); free(A);
}
long or_csr_nnz(const OrCsr *A) { return A->rowptr[A->nrows]; }

/* MatMult_SeqAIJ order: sum += a[k]*x[col[k]] over the row's stored entries, ascending column; the
 * product is rounded, then added (a stock x86-64 PETSc build has no FMA contraction). */
void or_csr_mult(const OrCsr *A, const double *x, double *y) {
#pragma omp parallel for schedule(static)
  for (int r = 0; r < A->nrows; ++r) {
    double s = 0.0;
    for (int k = A->rowptr[r]; k < A->rowptr[r + 1]; ++k) s += A->val[k] * x[A->col[k]];
    y[r] = s;
  }
}
void or_csr_mult_add(const OrCsr *A, const double *x, double *y) {
#pragma omp parallel for schedule(static)
  for (int r = 0; r < A->nrows; ++r) {
    double s = 0.0;
    for (int k = A->rowptr[r]; k < A->rowptr[r + 1]; ++k) s += A->val[k] * x[A->col[k]];
    y[r] += s;
  }
}
void or_csr_get_diagonal(const OrCsr *A, double *d) {
  for (int r = 0; r < A->nrows; ++r) {
    d[r] = 0.0;
    for (int k = A->rowptr[r]; k < A->rowptr[r + 1]; ++k)
      if (A->col[k] == r) { d[r] = A->val[k]; break; }
  }
}
OrCsr *or_csr_transpose(const OrCsr *A) {
  long nnz = or_csr_nnz(A);
  OrCsr *T = or_csr_alloc(A->ncols, A->nrows, nnz);
  for (long k = 0; k < nnz; ++k) T->rowptr[A->col[k] + 1]++;
  for (int c = 0; c < A->ncols; ++c) T->rowptr[c + 1] += T->rowptr[c];
  int *next = (int *)xmalloc(sizeof(int) * (size_t)A->ncols);
  memcpy(next, T->rowptr, sizeof(int) * (size_t)A->ncols);
  for (int r = 0; r < A->nrows; ++r)
    for (int k = A->rowptr[r]; k < A->rowptr[r + 1]; ++k) {
      int p = next[A->col[k]]++;
      T->col[p] = r; T->val[p] = A->val[k];
    }
  free(next);
  return T;
}
static int cmp_int(const void *a, const void *b) { int x = *(const int *)a, y = *(const int *)b; return (x > y) - (x < y); }
/* C = A*B.  Row-wise Gustavson; c_ij accumulates a_ik*b_kj in the order k appears in A's row
 * (MatMatMultNumeric_SeqAIJ_SeqAIJ order), each term a separately rounded product then add. */
OrCsr *or_csr_matmat(const OrCsr *A, const OrCsr *B) {
  int n = A->nrows, m = B->ncols;
  int *mark = (int *)xmalloc(sizeof(int) * (size_t)m);
  for (int j = 0; j < m; ++j) mark[j] = -1;
  int *rowptr = (int *)xcalloc((size_t)n + 1, sizeof(int));
  for (int i = 0; i < n; ++i) { /* symbolic */
    int cnt = 0;
    for (int ka = A->rowptr[i]; ka < A->rowptr[i + 1]; ++ka) {
      int k = A->col[ka];
      for (int kb = B->rowptr[k]; kb < B->rowptr[k + 1]; ++kb)
        if (mark[B->col[kb]] != i) { mark[B->col[kb]] = i; cnt++; }
    }
    rowptr[i + 1] = rowptr[i] + cnt;
  }
  OrCsr *C = or_csr_alloc(n, m, rowptr[n]);
  memcpy(C->rowptr, rowptr, sizeof(int) * ((size_t)n + 1));
  free(rowptr);
  for (int j = 0; j < m; ++j) mark[j] = -1;
  double *acc = (double *)xcalloc((size_t)m, sizeof(double));
  for (int i = 0; i < n; ++i) {
    int base = C->rowptr[i], cnt = 0;
    for (int ka = A->rowptr[i]; ka < A->rowptr[i + 1]; ++ka) {
      int k = A->col[ka];
      for (int kb = B->rowptr[k]; kb < B->rowptr[k + 1]; ++kb) {
        int j = B->col[kb];
        if (mark[j] != i) { mark[j] = i; C->col[base + cnt++] = j; acc[j] = 0.0; }
      }
    }
    qsort(C->col + base, (size_t)cnt, sizeof(int), cmp_int);
    for (int ka = A->rowptr[i]; ka < A->rowptr[i + 1]; ++ka) {
      int k = A->col[ka];
      double a = A->val[ka];
      for (int kb = B->rowptr[k]; kb < B->rowptr[k + 1]; ++kb) acc[B->col[kb]] += a * B->val[kb];
    }
    for (int t = 0; t < cnt; ++t) C->val[base + t] = acc[C->col[base + t]];
  }
  free(acc); free(mark);
  return C;
}
OrCsr *or_csr_add_scaled(const OrCsr *A, double a, const OrCsr *B) {
  int n = A->nrows;
  long cap = or_csr_nnz(A) + or_csr_nnz(B);
  OrCsr *C = or_csr_alloc(n, A->ncols, cap);
  long p = 0;
  for (int i = 0; i < n; ++i) {
    int ka = A->rowptr[i], ea = A->rowptr[i + 1], kb = B->rowptr[i], eb = B->rowptr[i + 1];
    while (ka < ea || kb < eb) {
      int ca = ka < ea ? A->col[ka] : 0x7fffffff, cb = kb < eb ? B->col[kb] : 0x7fffffff;
      if (ca == cb) { C->col[p] = ca; C->val[p++] = A->val[ka++] + a * B->val[kb++]; }
      else if (ca < cb) { C->col[p] = ca; C->val[p++] = A->val[ka++]; }
      else { C->col[p] = cb; C->val[p++] = a * B->val[kb++]; }
    }
    C->rowptr[i + 1] = (int)p;
  }
  return C;
}
OrCsr *or_csr_scale_cols(const OrCsr *A, const double *d) {
  long nnz = or_csr_nnz(A);
  OrCsr *C = or_csr_alloc(A->nrows, A->ncols, nnz);
  memcpy(C->rowptr, A->rowptr, sizeof(int) * ((size_t)A->nrows + 1));
  memcpy(C->col, A->col, sizeof(int) * (size_t)nnz);
  for (long k = 0; k < nnz; ++k) C->val[k] = A->val[k] * d[A->col[k]];
  return C;
}
void or_zero_rows(OrCsr *A, int n, const int *rows) {
  for (int t = 0; t < n; ++t)
    for (int k = A->rowptr[rows[t]]; k < A->rowptr[rows[t] + 1]; ++k) A->val[k] = 0.0;
}
void or_zero_cols(OrCsr *A, int n, const int *cols) {
  char *is = (char *)xcalloc((size_t)A->ncols, 1);
  for (int t = 0; t < n; ++t) is[cols[t]] = 1;
  long nnz = or_csr_nnz(A);
  for (long k = 0; k < nnz; ++k)
    if (is[A->col[k]]) A->val[k] = 0.0;
  free(is);
}

/* =============================================================== DMDA ===== */
/* DMSetUp_DA_2D with PETSC_DECIDE */
void or_dmda_proc_grid(int M, int N, int size, int *pm, int *pn) {
  int m = (int)(0.5 + sqrt(((double)M) * ((double)size) / ((double)N))), n = 1;
  if (!m) m = 1;
  while (m > 0) {
    n = size / m;
    if (m * n == size) break;
    m--;
  }
  if (M > N && m < n) { int t = m; m = n; n = t; }
  *pm = m; *pn = n;
}
void or_dmda_ownership(int M, int m, int *lx) {
  for (int i = 0; i < m; ++i) lx[i] = M / m + ((M % m) > i);
}
void or_dmda_natural_to_petsc(int M, int N, int size, int *node_map, int *node_owner) {
  int m, n;
  or_dmda_proc_grid(M, N, size, &m, &n);
  int *lx = (int *)xmalloc(sizeof(int) * (size_t)m), *ly = (int *)xmalloc(sizeof(int) * (size_t)n);
  or_dmda_ownership(M, m, lx);
  or_dmda_ownership(N, n, ly);
  int start = 0;
  for (int r = 0; r < size; ++r) {
    int pi = r % m, pj = r / m, xs = 0, ys = 0;
    for (int i = 0; i < pi; ++i) xs += lx[i];
    for (int j = 0; j < pj; ++j) ys += ly[j];
    for (int j = 0; j < ly[pj]; ++j)
      for (int i = 0; i < lx[pi]; ++i) {
        int nat = (ys + j) * M + xs + i;
        node_map[nat] = start + j * lx[pi] + i;
        if (node_owner) node_owner[nat] = r;
      }
    start += lx[pi] * ly[pj];
  }
  free(lx); free(ly);
}
/* DMDAGetElementsCorners / DMDAGetElementsSizes (Discretization.c:144-145) */
void or_dmda_element_range(int M, int N, int size, int rank, int *si, int *sj, int *ni, int *nj) {
  int m, n;
  or_dmda_proc_grid(M, N, size, &m, &n);
  int *lx = (int *)xmalloc(sizeof(int) * (size_t)m), *ly = (int *)xmalloc(sizeof(int) * (size_t)n);
  or_dmda_ownership(M, m, lx);
  or_dmda_ownership(N, n, ly);
  int pi = rank % m, pj = rank / m, xs = 0, ys = 0;
  for (int i = 0; i < pi; ++i) xs += lx[i];
  for (int j = 0; j < pj; ++j) ys += ly[j];
  int gxs = xs > 0 ? xs - 1 : xs, gys = ys > 0 ? ys - 1 : ys;
  *si = gxs; *sj = gys;
  *ni = xs + lx[pi] - gxs - 1;
  *nj = ys + ly[pj] - gys - 1;
  free(lx); free(ly);
}

/* ==================================================== element kernels ===== */
/* ConstructGaussQuadratureQ12D, Discretization.c:49-63 (truncated literal on purpose) */
static const double GP_XI[4][2] = {{-0.57735026919, -0.57735026919}, {-0.57735026919, 0.57735026919},
                                   {0.57735026919, 0.57735026919},   {0.57735026919, -0.57735026919}};
static const double GP_W[4] = {1.0, 1.0, 1.0, 1.0};

/* ConstructQ12D_Ni, Discretization.c:65-76 */
static void q1_Ni(const double xi_[2], double Ni[4]) {
  double xi = xi_[0], eta = xi_[1];
  Ni[0] = 0.25 * (1.0 - xi) * (1.0 - eta);
  Ni[1] = 0.25 * (1.0 - xi) * (1.0 + eta);
  Ni[2] = 0.25 * (1.0 + xi) * (1.0 + eta);
  Ni[3] = 0.25 * (1.0 + xi) * (1.0 - eta);
}
/* ConstructQ12D_GNi, Discretization.c:78-94 */
static void q1_GNi(const double xi_[2], double GNi[2][4]) {
  double xi = xi_[0], eta = xi_[1];
  GNi[0][0] = -0.25 * (1.0 - eta);
  GNi[0][1] = -0.25 * (1.0 + eta);
  GNi[0][2] = 0.25 * (1.0 + eta);
  GNi[0][3] = 0.25 * (1.0 - eta);
  GNi[1][0] = -0.25 * (1.0 - xi);
  GNi[1][1] = 0.25 * (1.0 - xi);
  GNi[1][2] = 0.25 * (1.0 + xi);
  GNi[1][3] = -0.25 * (1.0 + xi);
}
/* ConstructQ12D_GNx, Discretization.c:96-128 */
static void q1_GNx(double GNi[2][4], const double *ec, double GNx[2][4], double *detJ) {
  double Jac[2][2], invJ[2][2], J;
  for (int c = 0; c < 2; ++c)
    for (int d = 0; d < 2; ++d) Jac[c][d] = 0.0;
  for (int c = 0; c < 2; ++c)
    for (int d = 0; d < 2; ++d)
      for (int i = 0; i < 4; ++i) Jac[c][d] += GNi[c][i] * ec[i * 2 + d];
  J = Jac[0][0] * Jac[1][1] - Jac[0][1] * Jac[1][0];
  invJ[0][0] = Jac[1][1] / J;
  invJ[0][1] = -Jac[0][1] / J;
  invJ[1][0] = -Jac[1][0] / J;
  invJ[1][1] = Jac[0][0] / J;
  for (int i = 0; i < 4; ++i) {
    GNx[0][i] = invJ[0][0] * GNi[0][i] + invJ[0][1] * GNi[1][i];
    GNx[1][i] = invJ[1][0] * GNi[0][i] + invJ[1][1] * GNi[1][i];
  }
  *detJ = J;
}
/* DMDASetUniformCoordinates(0,1,0,1) (Discretization.c:25): x_i = 0 + hx*i, hx = 1/(M-1).
 * GetElementCoords (Discretization.c:31-46): as_written=1 reproduces the live lines :34-38
 * (all four nodes = node (ei,ej)); as_written=0 is the commented intent :40-43. */
void or_element_coords(int M, int N, int ei, int ej, int as_written, double ec[8]) {
  double hx = (1.0 - 0.0) / (double)(M - 1), hy = (1.0 - 0.0) / (double)(N - 1);
  static const int di[4] = {0, 0, 1, 1}, dj[4] = {0, 1, 1, 0};
  for (int k = 0; k < 4; ++k) {
    int i = ei + (as_written ? 0 : di[k]), j = ej + (as_written ? 0 : dj[k]);
    ec[2 * k + 0] = 0.0 + hx * (double)i;
    ec[2 * k + 1] = 0.0 + hy * (double)j;
  }
}
/* FormStressOperatorQ12D, Discretization.c:293-332.  Ke is accumulated as Ke[i+8*j]. */
void or_element_stress(const double ec[8], const double coeff[4], double Ke[64]) {
  for (int p = 0; p < 4; ++p) {
    double GNi[2][4], GNx[2][4], detJ, B[3][8], tildeD[3];
    q1_GNi(GP_XI[p], GNi);
    q1_GNx(GNi, ec, GNx, &detJ);
    for (int i = 0; i < 4; ++i) {
      B[0][2 * i] = GNx[0][i]; B[0][2 * i + 1] = 0.0;
      B[1][2 * i] = 0.0;       B[1][2 * i + 1] = GNx[1][i];
      B[2][2 * i] = GNx[1][i]; B[2][2 * i + 1] = GNx[0][i];
    }
    tildeD[0] = 2.0 * GP_W[p] * detJ * coeff[p];
    tildeD[1] = 2.0 * GP_W[p] * detJ * coeff[p];
    tildeD[2] = GP_W[p] * detJ * coeff[p];
    for (int i = 0; i < 8; ++i)
      for (int j = 0; j < 8; ++j)
        for (int k = 0; k < 3; ++k) Ke[i + 8 * j] += B[k][i] * tildeD[k] * B[k][j];
  }
}
/* FormLaplaceRHSQ12D + FormRHS, Discretization.c:334-374, 397-402.
 * kind 0: the reference's constant body force (1,2) (evaluated at the reference coordinate, :362-365,
 *         which is harmless for a constant).
 * kind 1: (ours, for the KKT workloads) rotational force (2y-1, 1-2x) at the physical Gauss point. */
void or_element_rhs(const double ec[8], int kind, double Fe[8]) {
  for (int p = 0; p < 4; ++p) {
    double Ni[4], GNi[2][4], GNx[2][4], detJ, fac, f_p[2];
    q1_Ni(GP_XI[p], Ni);
    q1_GNi(GP_XI[p], GNi);
    q1_GNx(GNi, ec, GNx, &detJ);
    fac = GP_W[p] * detJ;
    if (kind == 0) { f_p[0] = 1.0; f_p[1] = 2.0; }
    else {
      double xp = 0.0, yp = 0.0;
      for (int i = 0; i < 4; ++i) { xp += Ni[i] * ec[2 * i]; yp += Ni[i] * ec[2 * i + 1]; }
      f_p[0] = 2.0 * yp - 1.0;
      f_p[1] = 1.0 - 2.0 * xp;
    }
    for (int i = 0; i < 4; ++i)
      for (int c = 0; c < 2; ++c) Fe[i * 2 + c] += fac * Ni[i] * f_p[c];
  }
}
/* KKT element blocks.  NOT in the reference (its B is a stub, Discretization.c:277-290); defined here
 * in the ex43 lineage the reference names (main.c:1), same quadrature / node order / fac = w*detJ:
 *   Ge[(2i+d)*4+j] -= fac*GNx[d][i]*Ni[j]          gradient (u rows, p cols);  divergence = Ge^T
 *   Ce[i*4+j]      -= fac*(Ni[i]*Ni[j] - 0.0625)   Dohrmann-Bochev stabilisation = the (2,2) block
 *   Qe[i*4+j]      -= fac*(Ni[i]*Ni[j])            minus pressure mass matrix (Schur "user" matrix) */
void or_element_kkt(const double ec[8], double Ge[32], double Ce[16], double Qe[16]) {
  for (int p = 0; p < 4; ++p) {
    double Ni[4], GNi[2][4], GNx[2][4], detJ, fac;
    q1_Ni(GP_XI[p], Ni);
    q1_GNi(GP_XI[p], GNi);
    q1_GNx(GNi, ec, GNx, &detJ);
    fac = GP_W[p] * detJ;
    for (int i = 0; i < 4; ++i)
      for (int d = 0; d < 2; ++d)
        for (int j = 0; j < 4; ++j) Ge[(2 * i + d) * 4 + j] -= fac * GNx[d][i] * Ni[j];
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) {
        Ce[i * 4 + j] -= fac * (Ni[i] * Ni[j] - 0.0625);
        Qe[i * 4 + j] -= fac * (Ni[i] * Ni[j]);
      }
  }
}

/* =========================================================== assembly ===== */
/* DMCreateMatrix on a box-stencil DMDA: rows dofr per node, cols dofc per node,
 * all nodes of the clipped 3x3 box, ascending natural index, explicit zeros. */
static OrCsr *box_pattern(int M, int N, int dofr, int dofc) {
  long nnz = (long)dofr * dofc * (3L * M - 2) * (3L * N - 2);
  OrCsr *A = or_csr_alloc(dofr * M * N, dofc * M * N, nnz);
  long p = 0;
  for (int j = 0; j < N; ++j)
    for (int i = 0; i < M; ++i)
      for (int c = 0; c < dofr; ++c) {
        for (int jj = (j > 0 ? j - 1 : 0); jj <= (j < N - 1 ? j + 1 : N - 1); ++jj)
          for (int ii = (i > 0 ? i - 1 : 0); ii <= (i < M - 1 ? i + 1 : M - 1); ++ii)
            for (int cc = 0; cc < dofc; ++cc) A->col[p++] = (jj * M + ii) * dofc + cc;
        A->rowptr[(j * M + i) * dofr + c + 1] = (int)p;
      }
  return A;
}
/* MatSetValues(ADD_VALUES) into a preallocated row: locate the slot, += */
static void add_value(OrCsr *A, int row, int col, double v) {
  int lo = A->rowptr[row], hi = A->rowptr[row + 1];
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (A->col[mid] < col) lo = mid + 1; else hi = mid;
  }
  if (lo >= A->rowptr[row + 1] || A->col[lo] != col) { fprintf(stderr, "sp_oracle: entry (%d,%d) outside preallocation\n", row, col); abort(); }
  A->val[lo] += v;
}
/* DMDAGetElementEqnums node order (Discretization.c:377-395) */
static void element_nodes(int M, int ei, int ej, int nd[4]) {
  nd[0] = ej * M + ei;
  nd[1] = (ej + 1) * M + ei;
  nd[2] = (ej + 1) * M + ei + 1;
  nd[3] = ej * M + ei + 1;
}
/* AssembleOperator_Laplace, Discretization.c:130-172: elements j-outer / i-inner, coeff=1,
 * MatSetValuesStencil reads Ae ROW-major (v[a*8+b]) although Ke was accumulated as Ke[i+8*j]. */
OrCsr *or_assemble_A(int M, int N, int as_written) {
  OrCsr *A = box_pattern(M, N, 2, 2);
  for (int ej = 0; ej < N - 1; ++ej)
    for (int ei = 0; ei < M - 1; ++ei) {
      double ec[8], coeff[4] = {1.0, 1.0, 1.0, 1.0}, Ae[64];
      int nd[4];
      or_element_coords(M, N, ei, ej, as_written, ec);
      memset(Ae, 0, sizeof(Ae));
      or_element_stress(ec, coeff, Ae);
      element_nodes(M, ei, ej, nd);
      for (int a = 0; a < 8; ++a)
        for (int b = 0; b < 8; ++b) add_value(A, nd[a >> 1] * 2 + (a & 1), nd[b >> 1] * 2 + (b & 1), Ae[a * 8 + b]);
    }
  return A;
}
/* The coefficient is an INPUT of FormStressOperatorQ12D (one value per Gauss point, Discretization.c:151-157 sets 1.0).
 * kind 1 (ours, for the variable-coefficient SpMV measurements): a smooth viscosity 1 + x (1 - y) / 2 evaluated at the
 * physical Gauss point -- every stored value becomes distinct, so the tile dictionaries of the SpMV decline. */
OrCsr *or_assemble_A_coeff(int M, int N, int kind) {
  OrCsr *A = box_pattern(M, N, 2, 2);
  for (int ej = 0; ej < N - 1; ++ej)
    for (int ei = 0; ei < M - 1; ++ei) {
      double ec[8], coeff[4], Ae[64];
      int nd[4];
      or_element_coords(M, N, ei, ej, 0, ec);
      for (int p = 0; p < 4; ++p) {
        double Ni[4], xp = 0.0, yp = 0.0;
        q1_Ni(GP_XI[p], Ni);
        for (int i = 0; i < 4; ++i) { xp += Ni[i] * ec[2 * i]; yp += Ni[i] * ec[2 * i + 1]; }
        coeff[p] = kind ? 1.0 + 0.5 * xp * (1.0 - yp) : 1.0;
      }
      memset(Ae, 0, sizeof(Ae));
      or_element_stress(ec, coeff, Ae);
      element_nodes(M, ei, ej, nd);
      for (int a = 0; a < 8; ++a)
        for (int b = 0; b < 8; ++b) add_value(A, nd[a >> 1] * 2 + (a & 1), nd[b >> 1] * 2 + (b & 1), Ae[a * 8 + b]);
    }
  return A;
}
/* AssembleRHS_Laplace, Discretization.c:174-227 */
void or_assemble_rhs(int M, int N, int as_written, int kind, double *f) {
  v_zero(2 * M * N, f);
  for (int ej = 0; ej < N - 1; ++ej)
    for (int ei = 0; ei < M - 1; ++ei) {
      double ec[8], Fe[8];
      int nd[4];
      or_element_coords(M, N, ei, ej, as_written, ec);
      memset(Fe, 0, sizeof(Fe));
      or_element_rhs(ec, kind, Fe);
      element_nodes(M, ei, ej, nd);
      for (int n = 0; n < 4; ++n) {
        f[nd[n] * 2 + 0] += Fe[2 * n + 0];
        f[nd[n] * 2 + 1] += Fe[2 * n + 1];
      }
    }
}
/* ApplyBC_Laplace id list, Discretization.c:246-262 (single rank: natural == PETSc ordering) */
int or_bc_ids(int M, int N, int dof, int *ids) {
  int n = 0;
  for (int j = 0; j < N; ++j)
    for (int i = 0; i < M; ++i)
      if (i == 0 || i == M - 1 || j == 0 || j == N - 1)
        for (int d = 0; d < dof; ++d) ids[n++] = (j * M + i) * dof + d;
  return n;
}
/* VecSetValues(f,ids,0) + MatZeroRowsColumns(A,ids,1.0,NULL,NULL), Discretization.c:264-268, Appendix A.4 */
void or_apply_bc(OrCsr *A, double *f, int nbc, const int *ids) {
  if (f) for (int t = 0; t < nbc; ++t) f[ids[t]] = 0.0;
  or_zero_rows(A, nbc, ids);
  or_zero_cols(A, nbc, ids);
  for (int t = 0; t < nbc; ++t) add_value(A, ids[t], ids[t], 1.0);
}
void or_assemble_kkt(int M, int N, OrCsr **pBt, OrCsr **pB, OrCsr **pC, OrCsr **pQ) {
  OrCsr *Bt = box_pattern(M, N, 2, 1), *B = box_pattern(M, N, 1, 2), *C = box_pattern(M, N, 1, 1), *Q = box_pattern(M, N, 1, 1);
  for (int ej = 0; ej < N - 1; ++ej)
    for (int ei = 0; ei < M - 1; ++ei) {
      double ec[8], Ge[32], Ce[16], Qe[16];
      int nd[4];
      or_element_coords(M, N, ei, ej, 0, ec);
      memset(Ge, 0, sizeof(Ge)); memset(Ce, 0, sizeof(Ce)); memset(Qe, 0, sizeof(Qe));
      or_element_kkt(ec, Ge, Ce, Qe);
      element_nodes(M, ei, ej, nd);
      for (int a = 0; a < 8; ++a)
        for (int j = 0; j < 4; ++j) {
          add_value(Bt, nd[a >> 1] * 2 + (a & 1), nd[j], Ge[a * 4 + j]);
          add_value(B, nd[j], nd[a >> 1] * 2 + (a & 1), Ge[a * 4 + j]);
        }
      for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
          add_value(C, nd[i], nd[j], Ce[i * 4 + j]);
          add_value(Q, nd[i], nd[j], Qe[i * 4 + j]);
        }
    }
  *pBt = Bt; *pB = B; *pC = C; *pQ = Q;
}
/* Q1 interpolation, coarse Mc x Nc nodes -> fine (2Mc-1) x (2Nc-1) nodes (DMCreateInterpolation_DA_2D_Q1
 * weights 1, 1/2, 1/4), dof-interleaved.  bc: zero rows of fine boundary dofs and cols of coarse
 * boundary dofs (pattern kept) so the coarse correction never touches Dirichlet values. */
/* The reference's own constraint block: "barycentre and volume constraints" (main.c:1), B is 4 x nCols
 * (SaddlePointProblem.c:48-49), g has 4 entries (:51-52); AssembleOperator_Constraints / AssembleRHS_Constraints are
 * empty stubs (Discretization.c:277-290), so the rows are DEFINED here, with the reference's quadrature, node order
 * and fac = w*detJ, as the four lowest moments of the displacement field u = (Ux,Uy) about the domain centre c:
 *   row 0:  int Ux                      (barycentre, x)
 *   row 1:  int Uy                      (barycentre, y)
 *   row 2:  int (x-cx) Ux + (y-cy) Uy   (dilation moment: the first-order volume change of the material about c;
 *                                        the plain  int div u  vanishes identically under the reference's all-round
 *                                        Dirichlet condition and would make the row redundant)
 *   row 3:  int (x-cx) Uy - (y-cy) Ux   (rotation moment)
 * Element vector Be[r*8 + 2a+d] (node a, component d), summed over the Gauss points from +0.0. */
void or_element_constraints(const double ec[8], double Be[32]) {
  for (int p = 0; p < 4; ++p) {
    double Ni[4], GNi[2][4], GNx[2][4], detJ, fac, xp = 0.0, yp = 0.0, rx, ry;
    q1_Ni(GP_XI[p], Ni);
    q1_GNi(GP_XI[p], GNi);
    q1_GNx(GNi, ec, GNx, &detJ);
    fac = GP_W[p] * detJ;
    for (int i = 0; i < 4; ++i) { xp += Ni[i] * ec[2 * i]; yp += Ni[i] * ec[2 * i + 1]; }
    rx = xp - 0.5; ry = yp - 0.5;
    for (int a = 0; a < 4; ++a) {
      const double w = fac * Ni[a];
      Be[0 * 8 + 2 * a] += w;
      Be[1 * 8 + 2 * a + 1] += w;
      Be[2 * 8 + 2 * a] += w * rx;
      Be[2 * 8 + 2 * a + 1] += w * ry;
      Be[3 * 8 + 2 * a] -= w * ry;
      Be[3 * 8 + 2 * a + 1] += w * rx;
    }
  }
}
/* AssembleOperator_Constraints: B (4 x 2MN, rows 0/1 hold only their own component's columns, rows 2/3 all columns,
 * ascending) assembled with ADD_VALUES in the reference's element order; Bt is its transpose (2MN x 4, 3 entries
 * per row).  Columns of Dirichlet dofs are zeroed by the caller like MatZeroRowsColumns would on the KKT matrix. */
void or_assemble_constraints(int M, int N, OrCsr **pB, OrCsr **pBt) {
  const int n = 2 * M * N, nn = M * N;
  OrCsr *B = or_csr_alloc(4, n, 2L * nn + 2L * n);
  B->rowptr[0] = 0; B->rowptr[1] = nn; B->rowptr[2] = 2 * nn; B->rowptr[3] = 2 * nn + n; B->rowptr[4] = 2 * nn + 2 * n;
  for (int i = 0; i < nn; ++i) { B->col[i] = 2 * i; B->col[nn + i] = 2 * i + 1; }
  for (int i = 0; i < n; ++i) { B->col[2 * nn + i] = i; B->col[2 * nn + n + i] = i; }
  for (long k = 0; k < B->rowptr[4]; ++k) B->val[k] = 0.0;
  for (int ej = 0; ej < N - 1; ++ej)
    for (int ei = 0; ei < M - 1; ++ei) {
      double ec[8], Be[32];
      int nd[4];
      or_element_coords(M, N, ei, ej, 0, ec);
      memset(Be, 0, sizeof(Be));
      or_element_constraints(ec, Be);
      element_nodes(M, ei, ej, nd);
      for (int a = 0; a < 4; ++a) {
        B->val[nd[a]] += Be[0 * 8 + 2 * a];
        B->val[nn + nd[a]] += Be[1 * 8 + 2 * a + 1];
        for (int d = 0; d < 2; ++d) {
          B->val[2 * nn + 2 * nd[a] + d] += Be[2 * 8 + 2 * a + d];
          B->val[2 * nn + n + 2 * nd[a] + d] += Be[3 * 8 + 2 * a + d];
        }
      }
    }
  *pB = B;
  if (pBt) *pBt = or_csr_transpose(B);
}

OrCsr *or_interp_q1(int Mc, int Nc, int dof, int bc) {
  int Mf = 2 * Mc - 1, Nf = 2 * Nc - 1;
  long nnz = 0;
  for (int j = 0; j < Nf; ++j)
    for (int i = 0; i < Mf; ++i) nnz += (long)dof * ((i & 1) ? 2 : 1) * ((j & 1) ? 2 : 1);
  OrCsr *P = or_csr_alloc(dof * Mf * Nf, dof * Mc * Nc, nnz);
  long p = 0;
  for (int j = 0; j < Nf; ++j)
    for (int i = 0; i < Mf; ++i)
      for (int c = 0; c < dof; ++c) {
        int fb = (i == 0 || i == Mf - 1 || j == 0 || j == Nf - 1);
        int ni = (i & 1) ? 2 : 1, nj = (j & 1) ? 2 : 1;
        for (int b = 0; b < nj; ++b)
          for (int a = 0; a < ni; ++a) {
            int ic = i / 2 + a, jc = j / 2 + b;
            int cb = (ic == 0 || ic == Mc - 1 || jc == 0 || jc == Nc - 1);
            double w = (ni == 2 ? 0.5 : 1.0) * (nj == 2 ? 0.5 : 1.0);
            if (bc && (fb || cb)) w = 0.0;
            P->col[p] = (jc * Mc + ic) * dof + c;
            P->val[p++] = w;
          }
        P->rowptr[(j * Mf + i) * dof + c + 1] = (int)p;
      }
  return P;
}

/* ========================================================== operators ===== */
void or_op_apply(OrOp *op, const double *x, double *y) { op->apply(op, x, y); }
void or_op_free(OrOp *op) {
  if (!op) return;
  if (op->destroy) op->destroy(op);
  free(op);
}
static OrOp *op_new(int n_in, int n_out, void (*apply)(OrOp *, const double *, double *), void (*destroy)(OrOp *), void *ctx) {
  OrOp *o = (OrOp *)xmalloc(sizeof(OrOp));
  o->n_in = n_in; o->n_out = n_out; o->apply = apply; o->destroy = destroy; o->ctx = ctx;
  return o;
}
static void free_ctx(OrOp *o) { free(o->ctx); }

static void csr_apply(OrOp *o, const double *x, double *y) { or_csr_mult((const OrCsr *)o->ctx, x, y); }
OrOp *or_op_csr(const OrCsr *A) { return op_new(A->ncols, A->nrows, csr_apply, NULL, (void *)A); }

typedef struct { int n; double *d; } DiagCtx;
static void diag_apply(OrOp *o, const double *x, double *y) {
  DiagCtx *c = (DiagCtx *)o->ctx;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < c->n; ++i) y[i] = x[i] * c->d[i]; /* d holds reciprocals: VecPointwiseMult(y,x,1/diag) */
}
static void diag_destroy(OrOp *o) { DiagCtx *c = (DiagCtx *)o->ctx; free(c->d); free(c); }
OrOp *or_op_diag_inverse(int n, const double *d) {
  DiagCtx *c = (DiagCtx *)xmalloc(sizeof(DiagCtx));
  c->n = n; c->d = (double *)xmalloc(sizeof(double) * (size_t)n);
  for (int i = 0; i < n; ++i) c->d[i] = 1.0 / d[i]; /* VecReciprocal */
  return op_new(n, n, diag_apply, diag_destroy, c);
}
/* PCJACOBI: y = x ./ diag(P); a zero diagonal entry is replaced by 1 */
OrOp *or_op_jacobi(const OrCsr *A) {
  double *d = (double *)xmalloc(sizeof(double) * (size_t)A->nrows);
  or_csr_get_diagonal(A, d);
  for (int i = 0; i < A->nrows; ++i) if (d[i] == 0.0) d[i] = 1.0;
  OrOp *o = or_op_diag_inverse(A->nrows, d);
  free(d);
  return o;
}

typedef struct { const OrCsr *A00, *A01, *A10, *A11; } NestCtx;
static void nest_apply(OrOp *o, const double *x, double *y) {
  NestCtx *c = (NestCtx *)o->ctx;
  int n0 = c->A00->nrows;
  or_csr_mult(c->A00, x, y);
  or_csr_mult_add(c->A01, x + n0, y);
  or_csr_mult(c->A10, x, y + n0);
  if (c->A11) or_csr_mult_add(c->A11, x + n0, y + n0);
}
OrOp *or_op_nest(const OrCsr *A00, const OrCsr *A01, const OrCsr *A10, const OrCsr *A11) {
  NestCtx *c = (NestCtx *)xmalloc(sizeof(NestCtx));
  c->A00 = A00; c->A01 = A01; c->A10 = A10; c->A11 = A11;
  int n = A00->nrows + A10->nrows;
  return op_new(n, n, nest_apply, free_ctx, c);
}

/* MatSchurComplement: S x = A11 x - A10 ksp(A00) A01 x */
typedef struct { const OrCsr *A11, *A10, *A01; OrOp *K0; double *t0, *t1; } SchurCtx;
static void schur_apply(OrOp *o, const double *x, double *y) {
  SchurCtx *c = (SchurCtx *)o->ctx;
  int n1 = c->A10->nrows;
  or_csr_mult(c->A01, x, c->t0);
  or_op_apply(c->K0, c->t0, c->t1);
  or_csr_mult(c->A10, c->t1, y);
  v_scale(n1, -1.0, y);
  if (c->A11) or_csr_mult_add(c->A11, x, y);
}
static void schur_destroy(OrOp *o) { SchurCtx *c = (SchurCtx *)o->ctx; free(c->t0); free(c->t1); free(c); }
OrOp *or_op_schur(const OrCsr *A11, const OrCsr *A10, OrOp *K0, const OrCsr *A01) {
  SchurCtx *c = (SchurCtx *)xmalloc(sizeof(SchurCtx));
  c->A11 = A11; c->A10 = A10; c->A01 = A01; c->K0 = K0;
  c->t0 = (double *)xmalloc(sizeof(double) * (size_t)A01->nrows);
  c->t1 = (double *)xmalloc(sizeof(double) * (size_t)A01->nrows);
  return op_new(A10->nrows, A10->nrows, schur_apply, schur_destroy, c);
}

/* PCApply_FieldSplit_Schur */
typedef struct { int fact; const OrCsr *A01, *A10; OrOp *K0, *KS; double scale; double *t0, *t1; } FsCtx;
static void fs_apply(OrOp *o, const double *b, double *y) {
  FsCtx *c = (FsCtx *)o->ctx;
  int n0 = c->A01->nrows, n1 = c->A10->nrows;
  const double *b0 = b, *b1 = b + n0;
  double *y0 = y, *y1 = y + n0;
  switch (c->fact) {
  case 0: /* DIAG */
    or_op_apply(c->K0, b0, y0);
    or_op_apply(c->KS, b1, y1);
    v_scale(n1, c->scale, y1);
    break;
  case 1: /* LOWER */
    or_op_apply(c->K0, b0, y0);
    or_csr_mult(c->A10, y0, c->t1);
    v_aypx(n1, -1.0, b1, c->t1); /* t1 = b1 - A10 y0 */
    or_op_apply(c->KS, c->t1, y1);
    break;
  case 2: /* UPPER */
    or_op_apply(c->KS, b1, y1);
    or_csr_mult(c->A01, y1, c->t0);
    v_aypx(n0, -1.0, b0, c->t0); /* t0 = b0 - A01 y1 */
    or_op_apply(c->K0, c->t0, y0);
    break;
  default: /* FULL */
    or_op_apply(c->K0, b0, y0);
    or_csr_mult(c->A10, y0, c->t1);
    v_aypx(n1, -1.0, b1, c->t1);
    or_op_apply(c->KS, c->t1, y1);
    or_csr_mult(c->A01, y1, c->t0);
    v_aypx(n0, -1.0, b0, c->t0);
    or_op_apply(c->K0, c->t0, y0);
    break;
  }
}
static void fs_destroy(OrOp *o) { FsCtx *c = (FsCtx *)o->ctx; free(c->t0); free(c->t1); free(c); }
OrOp *or_op_fieldsplit(int fact, const OrCsr *A01, const OrCsr *A10, OrOp *K0, OrOp *KS, double scale) {
  FsCtx *c = (FsCtx *)xmalloc(sizeof(FsCtx));
  c->fact = fact; c->A01 = A01; c->A10 = A10; c->K0 = K0; c->KS = KS; c->scale = scale;
  c->t0 = (double *)xmalloc(sizeof(double) * (size_t)A01->nrows);
  c->t1 = (double *)xmalloc(sizeof(double) * (size_t)A10->nrows);
  int n = A01->nrows + A10->nrows;
  return op_new(n, n, fs_apply, fs_destroy, c);
}

/* PCLSC apply: y = Linv A10 [D^-1] A00 [D^-1] A01 Linv x */
typedef struct { const OrCsr *A00, *A01, *A10; OrOp *Linv; double *dinv; double *p0, *p1, *u0, *u1; } LscCtx;
static void lsc_apply(OrOp *o, const double *x, double *y) {
  LscCtx *c = (LscCtx *)o->ctx;
  int n0 = c->A00->nrows;
  or_op_apply(c->Linv, x, c->p0);
  or_csr_mult(c->A01, c->p0, c->u0);
  if (c->dinv) for (int i = 0; i < n0; ++i) c->u0[i] *= c->dinv[i];
  or_csr_mult(c->A00, c->u0, c->u1);
  if (c->dinv) for (int i = 0; i < n0; ++i) c->u1[i] *= c->dinv[i];
  or_csr_mult(c->A10, c->u1, c->p1);
  or_op_apply(c->Linv, c->p1, y);
}
static void lsc_destroy(OrOp *o) {
  LscCtx *c = (LscCtx *)o->ctx;
  free(c->dinv); free(c->p0); free(c->p1); free(c->u0); free(c->u1); free(c);
}
OrOp *or_op_lsc(const OrCsr *A00, const OrCsr *A01, const OrCsr *A10, OrOp *Linv, int scale_diag) {
  LscCtx *c = (LscCtx *)xmalloc(sizeof(LscCtx));
  int n0 = A00->nrows, n1 = A10->nrows;
  c->A00 = A00; c->A01 = A01; c->A10 = A10; c->Linv = Linv; c->dinv = NULL;
  if (scale_diag) {
    c->dinv = (double *)xmalloc(sizeof(double) * (size_t)n0);
    or_csr_get_diagonal(A00, c->dinv);
    for (int i = 0; i < n0; ++i) c->dinv[i] = 1.0 / c->dinv[i];
  }
  c->p0 = (double *)xmalloc(sizeof(double) * (size_t)n1); c->p1 = (double *)xmalloc(sizeof(double) * (size_t)n1);
  c->u0 = (double *)xmalloc(sizeof(double) * (size_t)n0); c->u1 = (double *)xmalloc(sizeof(double) * (size_t)n0);
  return op_new(n1, n1, lsc_apply, lsc_destroy, c);
}

/* fieldsplit on a monolithic, field-interleaved vector: gather the splits (VecScatter), apply, scatter back */
typedef struct { OrOp *inner; int n; int *map; double *xs, *ys; } PermCtx;
static void perm_apply(OrOp *o, const double *x, double *y) {
  PermCtx *c = (PermCtx *)o->ctx;
  for (int i = 0; i < c->n; ++i) c->xs[i] = x[c->map[i]];
  or_op_apply(c->inner, c->xs, c->ys);
  for (int i = 0; i < c->n; ++i) y[c->map[i]] = c->ys[i];
}
static void perm_destroy(OrOp *o) { PermCtx *c = (PermCtx *)o->ctx; free(c->map); free(c->xs); free(c->ys); free(c); }
OrOp *or_op_permuted(OrOp *inner, int n, const int *map) {
  PermCtx *c = (PermCtx *)xmalloc(sizeof(PermCtx));
  c->inner = inner; c->n = n;
  c->map = (int *)xmalloc(sizeof(int) * (size_t)n);
  memcpy(c->map, map, sizeof(int) * (size_t)n);
  c->xs = (double *)xmalloc(sizeof(double) * (size_t)n); c->ys = (double *)xmalloc(sizeof(double) * (size_t)n);
  return op_new(n, n, perm_apply, perm_destroy, c);
}

/* dense LU with partial pivoting (PCLU stand-in for the coarsest multigrid level) */
typedef struct { int n; double *lu; int *piv; } LuCtx;
static void lu_apply(OrOp *o, const double *b, double *x) {
  LuCtx *c = (LuCtx *)o->ctx;
  int n = c->n;
  for (int i = 0; i < n; ++i) x[i] = b[c->piv[i]];
  for (int i = 0; i < n; ++i) { double s = x[i]; for (int j = 0; j < i; ++j) s -= c->lu[(size_t)i * n + j] * x[j]; x[i] = s; }
  for (int i = n - 1; i >= 0; --i) { double s = x[i]; for (int j = i + 1; j < n; ++j) s -= c->lu[(size_t)i * n + j] * x[j]; x[i] = s / c->lu[(size_t)i * n + i]; }
}
static void lu_destroy(OrOp *o) { LuCtx *c = (LuCtx *)o->ctx; free(c->lu); free(c->piv); free(c); }
OrOp *or_op_dense_lu(const OrCsr *A) {
  int n = A->nrows;
  LuCtx *c = (LuCtx *)xmalloc(sizeof(LuCtx));
  c->n = n; c->lu = (double *)xcalloc((size_t)n * n, sizeof(double)); c->piv = (int *)xmalloc(sizeof(int) * (size_t)n);
  for (int r = 0; r < n; ++r)
    for (int k = A->rowptr[r]; k < A->rowptr[r + 1]; ++k) c->lu[(size_t)r * n + A->col[k]] = A->val[k];
  for (int i = 0; i < n; ++i) c->piv[i] = i;
  for (int k = 0; k < n; ++k) {
    int p = k; double best = fabs(c->lu[(size_t)k * n + k]);
    for (int i = k + 1; i < n; ++i) if (fabs(c->lu[(size_t)i * n + k]) > best) { best = fabs(c->lu[(size_t)i * n + k]); p = i; }
    if (p != k) {
      for (int j = 0; j < n; ++j) { double t = c->lu[(size_t)k * n + j]; c->lu[(size_t)k * n + j] = c->lu[(size_t)p * n + j]; c->lu[(size_t)p * n + j] = t; }
      int t = c->piv[k]; c->piv[k] = c->piv[p]; c->piv[p] = t;
    }
    for (int i = k + 1; i < n; ++i) {
      double l = c->lu[(size_t)i * n + k] / c->lu[(size_t)k * n + k];
      c->lu[(size_t)i * n + k] = l;
      for (int j = k + 1; j < n; ++j) c->lu[(size_t)i * n + j] -= l * c->lu[(size_t)k * n + j];
    }
  }
  return op_new(n, n, lu_apply, lu_destroy, c);
}

/* ================================================================ KSP ===== */
OrKsp *or_ksp_create(int type, OrOp *A, OrOp *M) {
  OrKsp *k = (OrKsp *)xcalloc(1, sizeof(OrKsp));
  k->type = type; k->A = A; k->M = M;
  k->rtol = 1e-5; k->atol = 1e-50; k->dtol = 1e5; k->max_it = 10000; k->restart = 30; /* PETSc defaults */
  k->richardson_scale = 1.0;
  return k;
}
void or_ksp_free(OrKsp *k) { if (k) { free(k->hist); free(k); } }
void or_ksp_set_history(OrKsp *k, int cap) {
  free(k->hist);
  k->hist = (double *)xmalloc(sizeof(double) * (size_t)cap);
  k->hist_cap = cap; k->hist_len = 0;
}
static void pc_apply(OrKsp *k, const double *x, double *y) {
  if (k->M) or_op_apply(k->M, x, y); else v_copy(k->A->n_in, x, y);
}
/* KSPConvergedDefault. returns reason (0 = keep iterating) */
static int converged(OrKsp *k, int it, double rnorm) {
  if (k->hist && k->hist_len < k->hist_cap) k->hist[k->hist_len++] = rnorm;
  k->rnorm = rnorm;
  if (it == 0) k->rnorm0 = rnorm;
  if (v_bad(rnorm)) return OR_DIVERGED_NANORINF;
  double ttol = fmax(k->rtol * k->rnorm0, k->atol);
  if (rnorm <= ttol) return rnorm < k->atol ? OR_CONVERGED_ATOL : OR_CONVERGED_RTOL;
  if (rnorm >= k->dtol * k->rnorm0) return OR_DIVERGED_DTOL;
  return 0;
}

/* KSPSolve_Richardson (no self-scale): x += scale * M^-1 (b - A x), max_it sweeps */
static int solve_richardson(OrKsp *k, const double *b, double *x, int guess_nonzero) {
  int n = k->A->n_in;
  double *r = (double *)xmalloc(sizeof(double) * (size_t)n), *z = (double *)xmalloc(sizeof(double) * (size_t)n);
  k->reason = 0;
  for (int it = 0; it < k->max_it; ++it) {
    if (it == 0 && !guess_nonzero) v_copy(n, b, r);
    else { or_op_apply(k->A, x, r); v_aypx(n, -1.0, b, r); }
    pc_apply(k, r, z);
    if (!k->norm_none) {
      k->reason = converged(k, it, or_norm2(n, z));
      if (k->reason) { k->its = it; break; }
    }
    v_axpy(n, k->richardson_scale, z, x);
    k->its = it + 1;
  }
  if (!k->reason) k->reason = k->norm_none ? OR_CONVERGED_ITS : OR_DIVERGED_ITS;
  free(r); free(z);
  return k->reason;
}

/* KSPSolve_Chebyshev recurrence; max_it PC applications */
static int solve_chebyshev(OrKsp *k, const double *b, double *x, int guess_nonzero) {
  int n = k->A->n_in;
  double *r = (double *)xmalloc(sizeof(double) * (size_t)n);
  double *p0 = (double *)xmalloc(sizeof(double) * (size_t)n), *p1 = (double *)xmalloc(sizeof(double) * (size_t)n), *p2 = (double *)xmalloc(sizeof(double) * (size_t)n);
  double *pkm1 = p0, *pk = p1, *pkp1 = p2;
  double scale = 2.0 / (k->emax + k->emin), alpha = 1.0 - scale * k->emin, mu = 1.0 / alpha, omegaprod = 2.0 / alpha;
  double ckm1 = 1.0, ck = mu, ckp1, omega;
  k->reason = 0;
  if (guess_nonzero) { or_op_apply(k->A, x, r); v_aypx(n, -1.0, b, r); } else v_copy(n, b, r);
  v_copy(n, x, pkm1);
  pc_apply(k, r, pk);                 /* pk = M^-1 r */
  if (!k->norm_none) k->reason = converged(k, 0, or_norm2(n, pk));
  v_aypx(n, scale, pkm1, pk);         /* pk = x + scale*z */
  k->its = 1;
  for (int i = 1; i < k->max_it && !k->reason; ++i) {
    or_op_apply(k->A, pk, r);
    v_aypx(n, -1.0, b, r);            /* r = b - A pk */
    ckp1 = 2.0 * mu * ck - ckm1;
    omega = omegaprod * ck / ckp1;
    pc_apply(k, r, pkp1);             /* z */
    if (!k->norm_none) { k->reason = converged(k, i, or_norm2(n, pkp1)); if (k->reason) break; }
    {
      double a = 1.0 - omega, bb = omega, c = omega * scale;
#pragma omp parallel for schedule(static)
      for (int t = 0; t < n; ++t) pkp1[t] = a * pkm1[t] + bb * pk[t] + c * pkp1[t]; /* VecAXPBYPCZ */
    }
    double *tp = pkm1; pkm1 = pk; pk = pkp1; pkp1 = tp;
    ckm1 = ck; ck = ckp1;
    k->its = i + 1;
  }
  v_copy(n, pk, x);
  if (!k->reason) k->reason = k->norm_none ? OR_CONVERGED_ITS : OR_DIVERGED_ITS;
  free(r); free(p0); free(p1); free(p2);
  return k->reason;
}

/* KSPSolve_GMRES / KSPSolve_FGMRES: classical Gram-Schmidt, no refinement.
 * GMRES: left PC, preconditioned residual norm.  FGMRES: right PC, true residual norm. */
static int solve_gmres(OrKsp *k, const double *b, double *x, int guess_nonzero, int flexible) {
  int n = k->A->n_in, m = k->restart;
  double *V = (double *)xmalloc(sizeof(double) * (size_t)n * (size_t)(m + 1));
  double *Z = flexible ? (double *)xmalloc(sizeof(double) * (size_t)n * (size_t)m) : NULL;
  double *w = (double *)xmalloc(sizeof(double) * (size_t)n), *t = (double *)xmalloc(sizeof(double) * (size_t)n);
  double *H = (double *)xcalloc((size_t)(m + 1) * (size_t)m, sizeof(double)); /* column-major: H[i + (m+1)*j] */
  double *cs = (double *)xmalloc(sizeof(double) * (size_t)m), *sn = (double *)xmalloc(sizeof(double) * (size_t)m);
  double *g = (double *)xmalloc(sizeof(double) * (size_t)(m + 1)), *y = (double *)xmalloc(sizeof(double) * (size_t)m);
  int its = 0, first = 1;
  k->reason = 0;
  while (!k->reason) {
    /* residual at cycle start */
    double *v0 = V;
    if (first && !guess_nonzero) {
      if (flexible) v_copy(n, b, v0); else pc_apply(k, b, v0);
    } else {
      or_op_apply(k->A, x, t);
      v_aypx(n, -1.0, b, t); /* t = b - A x */
      if (flexible) v_copy(n, t, v0); else pc_apply(k, t, v0);
    }
    first = 0;
    double beta = or_norm2(n, v0);
    k->reason = converged(k, its, beta);
    if (k->reason) break;
    if (its >= k->max_it) { k->reason = OR_DIVERGED_ITS; break; }
    v_scale(n, 1.0 / beta, v0);
    g[0] = beta;
    int it = 0;
    while (!k->reason && it < m && its < k->max_it) {
      double *vk = V + (size_t)n * it, *vn = V + (size_t)n * (it + 1);
      if (flexible) {
        double *zk = Z + (size_t)n * it;
        pc_apply(k, vk, zk);
        or_op_apply(k->A, zk, vn);
      } else {
        or_op_apply(k->A, vk, t);
        pc_apply(k, t, vn);
      }
      double *h = H + (size_t)(m + 1) * it;
      if (k->orthog == 3) { /* KSPGMRESModifiedGramSchmidtOrthogonalization */
        for (int j = 0; j <= it; ++j) { h[j] = or_dot(n, vn, V + (size_t)n * j); v_axpy(n, -h[j], V + (size_t)n * j, vn); }
      } else {               /* KSPGMRESClassicalGramSchmidtOrthogonalization */
        for (int j = 0; j <= it; ++j) h[j] = or_dot(n, vn, V + (size_t)n * j);         /* VecMDot */
        for (int j = 0; j <= it; ++j) v_axpy(n, -h[j], V + (size_t)n * j, vn);         /* VecMAXPY */
        int refine = k->orthog == 2;
        if (k->orthog == 1) { /* refine_ifneeded */
          double hnrm = 0.0, wn = or_dot(n, vn, vn);
          for (int j = 0; j <= it; ++j) hnrm += h[j] * h[j];
          refine = wn < hnrm;
        }
        if (refine) {
          double *h2 = (double *)xmalloc(sizeof(double) * (size_t)(it + 1));
          for (int j = 0; j <= it; ++j) h2[j] = or_dot(n, vn, V + (size_t)n * j);
          for (int j = 0; j <= it; ++j) v_axpy(n, -h2[j], V + (size_t)n * j, vn);
          for (int j = 0; j <= it; ++j) h[j] += h2[j];
          free(h2);
        }
      }
      double hn = or_norm2(n, vn);
      h[it + 1] = hn;
      if (hn != 0.0) v_scale(n, 1.0 / hn, vn);
      for (int j = 0; j < it; ++j) { /* previous rotations */
        double a = h[j], bb = h[j + 1];
        h[j] = cs[j] * a + sn[j] * bb;
        h[j + 1] = -sn[j] * a + cs[j] * bb;
      }
      double tt = sqrt(h[it] * h[it] + h[it + 1] * h[it + 1]);
      if (tt == 0.0) { k->reason = OR_DIVERGED_BREAKDOWN; break; }
      cs[it] = h[it] / tt; sn[it] = h[it + 1] / tt;
      g[it + 1] = -sn[it] * g[it];
      g[it] = cs[it] * g[it];
      h[it] = cs[it] * h[it] + sn[it] * h[it + 1];
      h[it + 1] = 0.0;
      double res = fabs(g[it + 1]);
      it++; its++;
      k->reason = converged(k, its, res);
      if (!k->reason && hn == 0.0) k->reason = OR_DIVERGED_BREAKDOWN;
    }
    /* BuildSolution: R y = g, x += V y (or Z y) */
    for (int i = it - 1; i >= 0; --i) {
      double s = g[i];
      for (int j = i + 1; j < it; ++j) s -= H[i + (size_t)(m + 1) * j] * y[j];
      y[i] = s / H[i + (size_t)(m + 1) * i];
    }
    for (int j = 0; j < it; ++j) v_axpy(n, y[j], (flexible ? Z : V) + (size_t)n * j, x);
    if (!k->reason && its >= k->max_it) k->reason = OR_DIVERGED_ITS;
  }
  k->its = its;
  free(V); free(Z); free(w); free(t); free(H); free(cs); free(sn); free(g); free(y);
  return k->reason;
}

/* KSPSolve_MINRES, classic Paige-Saunders form of PETSc <= 3.18 */
static int solve_minres(OrKsp *k, const double *b, double *x, int guess_nonzero) {
  int n = k->A->n_in;
  size_t bytes = sizeof(double) * (size_t)n;
  double *r = (double *)xmalloc(bytes), *z = (double *)xmalloc(bytes), *u = (double *)xmalloc(bytes), *v = (double *)xmalloc(bytes);
  double *uold = (double *)xcalloc((size_t)n, sizeof(double)), *vold = (double *)xcalloc((size_t)n, sizeof(double));
  double *w = (double *)xcalloc((size_t)n, sizeof(double)), *wold = (double *)xcalloc((size_t)n, sizeof(double)), *wooold = (double *)xmalloc(bytes);
  double alpha, beta, betaold = 1.0, eta, c = 1.0, cold = 1.0, s = 0.0, sold = 0.0, coold, soold, rho0, rho1, rho2, rho3, dp;
  int its = 0;
  k->reason = 0;
  if (guess_nonzero) { or_op_apply(k->A, x, r); v_aypx(n, -1.0, b, r); } else v_copy(n, b, r);
  pc_apply(k, r, z);
  dp = or_dot(n, r, z);
  if (dp < 0.0 && fabs(dp) > MINRES_HAPTOL) { k->reason = OR_DIVERGED_INDEFINITE_PC; goto done; }
  beta = sqrt(fabs(dp));
  eta = beta;
  { double rn = or_norm2(n, z); k->reason = converged(k, 0, rn); if (k->reason) goto done; dp = rn; }
  if (beta == 0.0) { k->reason = OR_CONVERGED_ATOL; goto done; }
  v_copy(n, r, v); v_copy(n, z, u);
  v_scale(n, 1.0 / beta, v); v_scale(n, 1.0 / beta, u);
  while (its < k->max_it) {
    or_op_apply(k->A, u, r);           /* r = A u */
    alpha = or_dot(n, u, r);
    pc_apply(k, r, z);                 /* z = M^-1 r */
    v_axpy(n, -alpha, v, r); v_axpy(n, -beta, vold, r);
    v_axpy(n, -alpha, u, z); v_axpy(n, -beta, uold, z);
    betaold = beta;
    { double d = or_dot(n, r, z); if (d < 0.0 && fabs(d) > MINRES_HAPTOL) { k->reason = OR_DIVERGED_INDEFINITE_PC; break; } beta = sqrt(fabs(d)); }
    coold = cold; cold = c; soold = sold; sold = s;
    rho0 = cold * alpha - coold * sold * betaold;
    rho1 = sqrt(rho0 * rho0 + beta * beta);
    rho2 = sold * alpha + coold * cold * betaold;
    rho3 = soold * betaold;
    c = rho0 / rho1; s = beta / rho1;
    v_copy(n, wold, wooold); v_copy(n, w, wold);
    {
      double irho1 = 1.0 / rho1;
#pragma omp parallel for schedule(static)
      for (int i = 0; i < n; ++i) w[i] = (u[i] - rho2 * wold[i] - rho3 * wooold[i]) * irho1; /* VecCopy,2xVecAXPY,VecScale */
    }
    v_axpy(n, c * eta, w, x);
    eta = -s * eta;
    v_copy(n, v, vold); v_copy(n, u, uold);
    v_copy(n, r, v); v_copy(n, z, u);
    if (beta != 0.0) { v_scale(n, 1.0 / beta, v); v_scale(n, 1.0 / beta, u); }
    dp = fabs(s) * dp;                 /* preconditioned residual norm estimate */
    its++;
    k->reason = converged(k, its, dp);
    if (k->reason) break;
  }
  if (!k->reason) k->reason = OR_DIVERGED_ITS;
done:
  k->its = its;
  free(r); free(z); free(u); free(v); free(uold); free(vold); free(w); free(wold); free(wooold);
  return k->reason;
}

int or_ksp_solve(OrKsp *k, const double *b, double *x, int guess_nonzero) {
  int n = k->A->n_in;
  k->its = 0; k->hist_len = 0;
  if (!guess_nonzero) v_zero(n, x);
  switch (k->type) {
  case OR_KSP_PREONLY:
    pc_apply(k, b, x);
    k->its = 1; k->reason = OR_CONVERGED_ITS;
    return k->reason;
  case OR_KSP_RICHARDSON: return solve_richardson(k, b, x, guess_nonzero);
  case OR_KSP_CHEBYSHEV: return solve_chebyshev(k, b, x, guess_nonzero);
  case OR_KSP_GMRES: return solve_gmres(k, b, x, guess_nonzero, 0);
  case OR_KSP_FGMRES: return solve_gmres(k, b, x, guess_nonzero, 1);
  case OR_KSP_MINRES: return solve_minres(k, b, x, guess_nonzero);
  }
  return OR_DIVERGED_BREAKDOWN;
}
static void kspop_apply(OrOp *o, const double *x, double *y) { or_ksp_solve((OrKsp *)o->ctx, x, y, 0); }
OrOp *or_op_from_ksp(OrKsp *k) { return op_new(k->A->n_in, k->A->n_in, kspop_apply, NULL, k); }

double or_estimate_lambda_max(OrOp *A, OrOp *M, int nits) {
  int n = A->n_in;
  double *v = (double *)xmalloc(sizeof(double) * (size_t)n), *t = (double *)xmalloc(sizeof(double) * (size_t)n), *z = (double *)xmalloc(sizeof(double) * (size_t)n);
  or_hash_vector(n, v);
  double lam = 0.0;
  for (int it = 0; it < nits; ++it) {
    double nv = or_norm2(n, v);
    v_scale(n, 1.0 / nv, v);
    or_op_apply(A, v, t);
    if (M) or_op_apply(M, t, z); else v_copy(n, t, z);
    lam = or_norm2(n, z);
    v_copy(n, z, v);
  }
  free(v); free(t); free(z);
  return lam;
}

/* ================================================================= MG ===== */
/* PCMG, multiplicative V(1,1)-cycle in PETSc's sense: smoothdown, residual, restrict (R = P^T),
 * recurse, interpolate-add, smoothup.  The smoothers are KSPs run with norm_none. */
typedef struct { int nlev; const OrCsr **A, **P; OrCsr **R; OrKsp **smooth; OrOp *coarse; double **b, **x, **r; } MgCtx;
static void mg_cycle(MgCtx *c, int l) {
  int n = c->A[l]->nrows;
  if (l == c->nlev - 1) { or_op_apply(c->coarse, c->b[l], c->x[l]); return; }
  or_ksp_solve(c->smooth[l], c->b[l], c->x[l], 0);          /* pre-smooth, zero initial guess */
  or_csr_mult(c->A[l], c->x[l], c->r[l]);
  v_aypx(n, -1.0, c->b[l], c->r[l]);                       /* r = b - A x */
  or_csr_mult(c->R[l], c->r[l], c->b[l + 1]);
  mg_cycle(c, l + 1);
  or_csr_mult_add(c->P[l], c->x[l + 1], c->x[l]);
  or_ksp_solve(c->smooth[l], c->b[l], c->x[l], 1);          /* post-smooth */
}
static void mg_apply(OrOp *o, const double *b, double *x) {
  MgCtx *c = (MgCtx *)o->ctx;
  int n = c->A[0]->nrows;
  v_copy(n, b, c->b[0]);
  mg_cycle(c, 0);
  v_copy(n, c->x[0], x);
}
static void mg_destroy(OrOp *o) {
  MgCtx *c = (MgCtx *)o->ctx;
  for (int l = 0; l < c->nlev; ++l) { free(c->b[l]); free(c->x[l]); free(c->r[l]); if (l < c->nlev - 1) or_csr_free(c->R[l]); }
  free(c->A); free(c->P); free(c->R); free(c->smooth); free(c->b); free(c->x); free(c->r); free(c);
}
OrOp *or_op_mg(int nlev, const OrCsr **A, const OrCsr **P, OrKsp **smooth, OrOp *coarse) {
  MgCtx *c = (MgCtx *)xmalloc(sizeof(MgCtx));
  c->nlev = nlev; c->coarse = coarse;
  c->A = (const OrCsr **)xmalloc(sizeof(void *) * (size_t)nlev);
  c->P = (const OrCsr **)xmalloc(sizeof(void *) * (size_t)nlev);
  c->R = (OrCsr **)xmalloc(sizeof(void *) * (size_t)nlev);
  c->smooth = (OrKsp **)xmalloc(sizeof(void *) * (size_t)nlev);
  c->b = (double **)xmalloc(sizeof(void *) * (size_t)nlev); c->x = (double **)xmalloc(sizeof(void *) * (size_t)nlev); c->r = (double **)xmalloc(sizeof(void *) * (size_t)nlev);
  for (int l = 0; l < nlev; ++l) {
    c->A[l] = A[l];
    size_t bytes = sizeof(double) * (size_t)A[l]->nrows;
    c->b[l] = (double *)xmalloc(bytes); c->x[l] = (double *)xmalloc(bytes); c->r[l] = (double *)xmalloc(bytes);
    if (l < nlev - 1) { c->P[l] = P[l]; c->R[l] = or_csr_transpose(P[l]); c->smooth[l] = smooth[l]; }
    else { c->P[l] = NULL; c->R[l] = NULL; c->smooth[l] = NULL; }
  }
  return op_new(A[0]->nrows, A[0]->nrows, mg_apply, mg_destroy, c);
}
