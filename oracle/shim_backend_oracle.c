/* shim_backend_oracle.c -- PETSc-shim back end over the CPU ORACLE.  TEST INFRASTRUCTURE ONLY: it exists so the
 * UNMODIFIED reference sources can be run on the CPU (oracle/_ref) to pin the oracle's restatement of
 * src/Discretization.c, and to run the reference's main.c as the CPU leg of config 0.  The product shim links
 * saddle_point_petsc_b200/csrc/shim_backend_b200sp.c instead. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "sp_oracle.h"
#include "../saddle_point_petsc_b200/csrc/shim_backend.h"

struct shimbk_mat_s { OrCsr *A; int M, N, dof; };
static char g_err[512];
static double *g_hist = NULL;
static int g_hist_len = 0;
const char *shimbk_name(void) { return "sp_oracle (CPU, test infrastructure)"; }
const char *shimbk_last_error(void) { return g_err; }
int shimbk_init(void) { return 0; }
int shimbk_finalize(void) { return 0; }

/* MatSetValues(ADD_VALUES) semantics restated independently of or_assemble_A: stable counting sort by (row, col),
 * duplicates summed in insertion order from +0.0 */
typedef struct { int r, c; long pos; } Trip;
static int trip_cmp(const void *a, const void *b) {
  const Trip *x = (const Trip *)a, *y = (const Trip *)b;
  if (x->r != y->r) return x->r < y->r ? -1 : 1;
  if (x->c != y->c) return x->c < y->c ? -1 : 1;
  return x->pos < y->pos ? -1 : (x->pos > y->pos ? 1 : 0);
}
int shimbk_mat_from_coo(int nrows, int ncols, long ncoo, const int *row, const int *col, const double *val, shimbk_mat *out) {
  Trip *t = (Trip *)malloc(sizeof(Trip) * (size_t)(ncoo ? ncoo : 1));
  for (long i = 0; i < ncoo; ++i) {
    if (row[i] < 0 || row[i] >= nrows || col[i] < 0 || col[i] >= ncols) { snprintf(g_err, sizeof(g_err), "COO index out of range"); free(t); return 1; }
    t[i].r = row[i]; t[i].c = col[i]; t[i].pos = i;
  }
  qsort(t, (size_t)ncoo, sizeof(Trip), trip_cmp);
  long nuniq = 0;
  for (long i = 0; i < ncoo; ++i) if (i == 0 || t[i].r != t[i - 1].r || t[i].c != t[i - 1].c) nuniq++;
  OrCsr *A = or_csr_alloc(nrows, ncols, nuniq);
  long u = -1;
  for (long i = 0; i < ncoo; ++i) {
    if (i == 0 || t[i].r != t[i - 1].r || t[i].c != t[i - 1].c) { ++u; A->col[u] = t[i].c; A->val[u] = 0.0; A->rowptr[t[i].r + 1]++; }
    A->val[u] += val[t[i].pos];
  }
  for (int r = 0; r < nrows; ++r) A->rowptr[r + 1] += A->rowptr[r];
  free(t);
  shimbk_mat h = (shimbk_mat)calloc(1, sizeof(*h));
  h->A = A;
  *out = h;
  return 0;
}
int shimbk_mat_set_grid(shimbk_mat A, int M, int N, int dof) { A->M = M; A->N = N; A->dof = dof; return 0; }
int shimbk_mat_zero_rows_columns(shimbk_mat A, int n, const int *rows, double diag) {
  or_zero_rows(A->A, n, rows);
  or_zero_cols(A->A, n, rows);
  for (int t = 0; t < n; ++t)
    for (int k = A->A->rowptr[rows[t]]; k < A->A->rowptr[rows[t] + 1]; ++k)
      if (A->A->col[k] == rows[t]) A->A->val[k] = diag;
  return 0;
}
int shimbk_mat_get_csr(shimbk_mat A, int *nrows, long *nnz, int *rowptr, int *col, double *val) {
  long nz = or_csr_nnz(A->A);
  if (nrows) *nrows = A->A->nrows;
  if (nnz) *nnz = nz;
  if (rowptr) memcpy(rowptr, A->A->rowptr, sizeof(int) * ((size_t)A->A->nrows + 1));
  if (col) memcpy(col, A->A->col, sizeof(int) * (size_t)nz);
  if (val) memcpy(val, A->A->val, sizeof(double) * (size_t)nz);
  return 0;
}
int shimbk_mat_destroy(shimbk_mat A) { if (A) { or_csr_free(A->A); free(A); } return 0; }

static const char *find_opt(const char *opts, const char *name, char *buf, size_t n) {
  char key[128];
  snprintf(key, sizeof(key), "-%s ", name);
  const char *p = strstr(opts, key);
  if (!p) return NULL;
  p += strlen(key);
  size_t k = 0;
  while (*p && *p != ' ' && k + 1 < n) buf[k++] = *p++;
  buf[k] = 0;
  return buf;
}
/* the velocity-block solves the reference's own driver can request: -ksp_type {gmres,fgmres,minres} -pc_type {none,jacobi} */
int shimbk_ksp_solve(shimbk_mat A, const char *options, int n, const double *b, double *x, int *its, int *reason, double *rnorm) {
  char buf[64];
  int type = OR_KSP_GMRES;
  const char *v = find_opt(options, "ksp_type", buf, sizeof(buf));
  if (v) type = !strcmp(v, "fgmres") ? OR_KSP_FGMRES : !strcmp(v, "minres") ? OR_KSP_MINRES : !strcmp(v, "gmres") ? OR_KSP_GMRES : -1;
  if (type < 0) { snprintf(g_err, sizeof(g_err), "oracle back end: unsupported -ksp_type %s", v); return 1; }
  OrOp *Aop = or_op_csr(A->A), *M = NULL;
  v = find_opt(options, "pc_type", buf, sizeof(buf));
  if (v && !strcmp(v, "jacobi")) M = or_op_jacobi(A->A);
  else if (v && strcmp(v, "none")) { snprintf(g_err, sizeof(g_err), "oracle back end: unsupported -pc_type %s", v); or_op_free(Aop); return 1; }
  OrKsp *k = or_ksp_create(type, Aop, M);
  if ((v = find_opt(options, "ksp_rtol", buf, sizeof(buf)))) k->rtol = atof(v);
  if ((v = find_opt(options, "ksp_max_it", buf, sizeof(buf)))) k->max_it = atoi(v);
  if ((v = find_opt(options, "ksp_gmres_restart", buf, sizeof(buf)))) k->restart = atoi(v);
  (void)n;
  or_ksp_set_history(k, k->max_it + 2 + k->max_it / (k->restart > 0 ? k->restart : 1));
  or_ksp_solve(k, b, x, 0);
  *its = k->its; *reason = k->reason; *rnorm = k->rnorm;
  free(g_hist);
  g_hist_len = k->hist_len;
  g_hist = (double *)malloc(sizeof(double) * (size_t)(g_hist_len > 0 ? g_hist_len : 1));
  for (int i = 0; i < g_hist_len; ++i) g_hist[i] = k->hist[i];
  or_ksp_free(k); or_op_free(Aop); if (M) or_op_free(M);
  return 0;
}
int shimbk_ksp_history(double *hist, int cap, int *len) {
  if (len) *len = g_hist_len;
  if (hist) for (int i = 0; i < g_hist_len && i < cap; ++i) hist[i] = g_hist[i];
  return 0;
}
