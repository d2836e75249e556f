/* shim_backend.h -- the operations the PETSc shim (petsc_shim.c) needs from a solver back end.
 * Selected at LINK time:
 *   saddle_point_petsc_b200/csrc/shim_backend_b200sp.c  -> libb200sp C ABI (the product, GPU only)
 *   oracle/shim_backend_oracle.c                        -> the CPU oracle (test infrastructure, oracle/_ref only)
 */
#ifndef B200SP_SHIM_BACKEND_H
#define B200SP_SHIM_BACKEND_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct shimbk_mat_s *shimbk_mat;

int shimbk_init(void);
int shimbk_finalize(void);
const char *shimbk_name(void);
const char *shimbk_last_error(void);
/* MatAssemblyEnd: COO triplets in MatSetValues call order -> sorted CSR, duplicates summed in that order */
int shimbk_mat_from_coo(int nrows, int ncols, long ncoo, const int *row, const int *col, const double *val, shimbk_mat *A);
/* tell the back end which DMDA grid (M x N nodes, dof) the matrix lives on (needed by -pc_type mg) */
int shimbk_mat_set_grid(shimbk_mat A, int M, int N, int dof);
int shimbk_mat_zero_rows_columns(shimbk_mat A, int n, const int *rows, double diag);
/* rowptr/col/val may be NULL to query nrows / nnz */
int shimbk_mat_get_csr(shimbk_mat A, int *nrows, long *nnz, int *rowptr, int *col, double *val);
int shimbk_mat_destroy(shimbk_mat A);
/* KSPSetFromOptions + KSPSetUp + KSPSolve with host vectors; options is PETSc options-database text */
int shimbk_ksp_solve(shimbk_mat A, const char *options, int n, const double *b, double *x, int *its, int *reason, double *rnorm);

/* residual history of the last shimbk_ksp_solve (what KSPMonitor would have been called with, in order; GMRES/FGMRES log
 * the recomputed residual again at the start of every restart cycle).  hist may be NULL to query the length. */
int shimbk_ksp_history(double *hist, int cap, int *len);

#ifdef __cplusplus
}
#endif
#endif
