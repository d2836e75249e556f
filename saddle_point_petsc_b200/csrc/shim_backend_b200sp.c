/* shim_backend_b200sp.c -- PETSc-shim back end over the libb200sp C ABI (the product path: GPU only). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/b200sp.h"
#include "shim_backend.h"

struct shimbk_mat_s { b200sp_mat m; };
static b200sp_ctx g_ctx = NULL;
static char g_err[1024];
static double *g_hist = NULL;
static int g_hist_len = 0;

#define CK(call) do { int rc_ = (call); if (rc_) { snprintf(g_err, sizeof(g_err), "%s -> %d: %s", #call, rc_, b200sp_last_error()); return rc_; } } while (0)

const char *shimbk_name(void) { return "libb200sp (CUDA sm_100a)"; }
const char *shimbk_last_error(void) { return g_err; }
int shimbk_init(void) {
  if (g_ctx) return 0;
  const char *dev = getenv("B200SP_DEVICE");
  CK(b200sp_ctx_create(dev ? atoi(dev) : 0, 0, 1, NULL, &g_ctx)); /* fails loudly without a B200: no CPU fallback */
  return 0;
}
int shimbk_finalize(void) {
  if (g_ctx) { CK(b200sp_ctx_destroy(g_ctx)); g_ctx = NULL; }
  return 0;
}
int shimbk_mat_from_coo(int nrows, int ncols, long ncoo, const int *row, const int *col, const double *val, shimbk_mat *A) {
  shimbk_mat h = (shimbk_mat)calloc(1, sizeof(*h));
  int rc = b200sp_mat_create_coo(g_ctx, nrows, ncols, ncoo, row, col, val, &h->m);
  if (rc) { snprintf(g_err, sizeof(g_err), "b200sp_mat_create_coo -> %d: %s", rc, b200sp_last_error()); free(h); return rc; }
  *A = h;
  return 0;
}
int shimbk_mat_set_grid(shimbk_mat A, int M, int N, int dof) { CK(b200sp_mat_set_grid(A->m, M, N, dof)); return 0; }
int shimbk_mat_zero_rows_columns(shimbk_mat A, int n, const int *rows, double diag) { CK(b200sp_mat_zero_rows_columns(A->m, n, rows, diag)); return 0; }
int shimbk_mat_get_csr(shimbk_mat A, int *nrows, long *nnz, int *rowptr, int *col, double *val) {
  int r, c; int64_t z;
  CK(b200sp_mat_get_size(A->m, &r, &c, &z));
  if (nrows) *nrows = r;
  if (nnz) *nnz = (long)z;
  if (rowptr || col || val) CK(b200sp_mat_get_csr_host(A->m, rowptr, col, val));
  return 0;
}
int shimbk_mat_destroy(shimbk_mat A) { if (A) { b200sp_mat_destroy(A->m); free(A); } return 0; }
int shimbk_ksp_solve(shimbk_mat A, const char *options, int n, const double *b, double *x, int *its, int *reason, double *rnorm) {
  b200sp_ksp ksp = NULL;
  CK(b200sp_ksp_create(g_ctx, &ksp));
  int rc = b200sp_ksp_set_operators(ksp, A->m, A->m);
  if (!rc) rc = b200sp_ksp_set_options(ksp, options);
  if (!rc) rc = b200sp_ksp_setup(ksp);
  if (!rc) rc = b200sp_ksp_solve_host(ksp, b, x, n);
  if (!rc) rc = b200sp_ksp_get_iteration_number(ksp, its);
  if (!rc) rc = b200sp_ksp_get_converged_reason(ksp, reason);
  if (!rc) rc = b200sp_ksp_get_residual_norm(ksp, rnorm);
  if (!rc) { /* keep the residual history for -ksp_monitor */
    int len = 0;
    rc = b200sp_ksp_get_residual_history(ksp, NULL, 0, &len);
    if (!rc) {
      free(g_hist);
      g_hist = (double *)malloc(sizeof(double) * (size_t)(len > 0 ? len : 1));
      g_hist_len = 0;
      rc = b200sp_ksp_get_residual_history(ksp, g_hist, len, &g_hist_len);
    }
  }
  if (rc) snprintf(g_err, sizeof(g_err), "KSP -> %d: %s", rc, b200sp_last_error());
  b200sp_ksp_destroy(&ksp);
  return rc;
}
int shimbk_ksp_history(double *hist, int cap, int *len) {
  if (len) *len = g_hist_len;
  if (hist) for (int i = 0; i < g_hist_len && i < cap; ++i) hist[i] = g_hist[i];
  return 0;
}
