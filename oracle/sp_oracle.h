/*
 * sp_oracle.h -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * Plain-C fp64 restatement of the hot path of p-m-mueller/saddle_point_petsc:
 *   - the reference's own Q1 discretisation (src/Discretization.c), restated
 *     in the reference's exact evaluation order, and
 *   - the PETSc algorithms the reference reaches through KSPSolve
 *     (src/SaddlePointProblem.c:65-72): CSR SpMV, Jacobi, Chebyshev, GMRES,
 *     FGMRES, MINRES, PCFIELDSPLIT/Schur, PCLSC, PCMG.
 *
 * PARITY STATUS: the element kernels and the A/f assembly + BC are PINNED
 * bit-for-bit against the real reference code compiled from /root/reference
 * (oracle/ref_build.sh -> oracle/_ref, golden vectors in tests/golden/,
 * tests/test_golden.py).  The Krylov /
 * preconditioner part is "parity unpinned": PETSc is an un-vendored,
 * un-pinned dependency of the reference (CMakeLists.txt:13) that is absent
 * from this image, and the reference ships no tests or golden vectors, so
 * that part restates PETSc's documented algorithms.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may link or call this code.
 */
#ifndef SP_ORACLE_H
#define SP_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OrCsr {
  int nrows, ncols;
  int *rowptr; /* nrows+1 */
  int *col;    /* nnz, ascending within a row */
  double *val; /* nnz */
} OrCsr;

/* ---- CSR containers ---- */
OrCsr *or_csr_alloc(int nrows, int ncols, long nnz);
void or_csr_free(OrCsr *A);
long or_csr_nnz(const OrCsr *A);
OrCsr *or_csr_transpose(const OrCsr *A);
OrCsr *or_csr_matmat(const OrCsr *A, const OrCsr *B);               /* C = A*B, Gustavson, sums in A-row order */
OrCsr *or_csr_add_scaled(const OrCsr *A, double a, const OrCsr *B); /* A + a*B on the union pattern */
OrCsr *or_csr_scale_cols(const OrCsr *A, const double *d);          /* A*diag(d) (copy) */
void or_csr_get_diagonal(const OrCsr *A, double *d);
void or_csr_mult(const OrCsr *A, const double *x, double *y);       /* sequential mul+add, CSR order */
void or_csr_mult_add(const OrCsr *A, const double *x, double *y);   /* y += A x */

/* ---- DMDA partition (PETSc DMDACreate2d semantics; src/Discretization.c:17) ---- */
void or_dmda_proc_grid(int M, int N, int size, int *pm, int *pn);
void or_dmda_ownership(int M, int m, int *lx); /* lx[m] */
/* natural node id (j*M+i) -> PETSc global node id for a size-rank run */
void or_dmda_natural_to_petsc(int M, int N, int size, int *node_map /* M*N */, int *node_owner /* M*N, may be NULL */);
void or_dmda_element_range(int M, int N, int size, int rank, int *si, int *sj, int *ni, int *nj);

/* ---- element kernels (src/Discretization.c:49-128, 293-374) ---- */
void or_element_coords(int M, int N, int ei, int ej, int as_written, double ec[8]);
void or_element_stress(const double ec[8], const double coeff[4], double Ke[64]);        /* += ; FormStressOperatorQ12D */
void or_element_rhs(const double ec[8], int kind, double Fe[8]);                         /* += ; FormLaplaceRHSQ12D */
void or_element_kkt(const double ec[8], double Ge[32], double Ce[16], double Qe[16]);    /* += ; KKT blocks (ours) */

/* ---- global assembly (src/Discretization.c:130-274), single-rank natural ordering ---- */
OrCsr *or_assemble_A_coeff(int M, int N, int kind);                  /* same with a variable coefficient per Gauss point (kind 1) */
OrCsr *or_assemble_A(int M, int N, int as_written);                  /* DMCreateMatrix + AssembleOperator_Laplace */
void or_assemble_rhs(int M, int N, int as_written, int kind, double *f /* 2*M*N */);
int or_bc_ids(int M, int N, int dof, int *ids /* dof*(2M+2N-4) */);  /* ApplyBC_Laplace ids, ascending */
void or_apply_bc(OrCsr *A, double *f, int nbc, const int *ids);      /* MatZeroRowsColumns(diag=1) + f=0 */
/* KKT blocks on the same nodal grid (pressure nodal, 1 dof/node):
 * Bt (2MN x MN) gradient, B = Bt^T, C (MN x MN) stabilisation (the (2,2) block, sign included),
 * Q (MN x MN) = -pressure mass matrix (the "user" Schur preconditioning matrix). */
void or_assemble_kkt(int M, int N, OrCsr **Bt, OrCsr **B, OrCsr **C, OrCsr **Q);
/* ---- 3-D Q1 hexahedron analogue (sp_oracle3d.c; BASELINE config 4 -- no reference code exists: DIM 2 is a #define) ---- */
void or3_element_coords(int M, int N, int P, int ei, int ej, int ek, double ec[24]);
void or3_element_stress(const double ec[24], double Ke[576]);                           /* += */
void or3_element_rhs(const double ec[24], int kind, double Fe[24]);                     /* += */
void or3_element_kkt(const double ec[24], double Ge[192], double Ce[64], double Qe[64]); /* += */
OrCsr *or3_assemble_A(int M, int N, int P);
void or3_assemble_rhs(int M, int N, int P, int kind, double *f);
void or3_assemble_kkt(int M, int N, int P, OrCsr **Bt, OrCsr **B, OrCsr **C, OrCsr **Q);
int or3_bc_ids(int M, int N, int P, int dof, int *ids); /* ids may be NULL to query the count */
void or3_dmda_proc_grid(int M, int N, int P, int size, int *pm, int *pn, int *pp);
void or3_dmda_natural_to_petsc(int M, int N, int P, int size, int *node_map, int *node_owner);
void or_element_constraints(const double ec[8], double Be[32]);               /* += ; the reference's 4 constraint rows (ours) */
void or_assemble_constraints(int M, int N, OrCsr **B, OrCsr **Bt);            /* B: 4 x 2MN, Bt = B^T */
void or_zero_rows(OrCsr *A, int n, const int *rows);
void or_zero_cols(OrCsr *A, int n, const int *cols);
/* ---- aggregation multigrid set-up (sp_oracle_amg.c; ours -- PCGAMG analogue, deterministic MIS-2 aggregation) ---- */
int or_amg_aggregate(const OrCsr *A, int bs, double theta, int order /* 0 hashed, 1 natural */, int *agg /* nrows/bs; -1 = left out */); /* returns #aggregates */
OrCsr *or_amg_tentative(int nn, int bs, const int *agg, int nagg, const int *w /* NULL: ones */, int *wc /* nagg, may be NULL */);
OrCsr *or_csr_scale_rows(const OrCsr *A, const double *d);                    /* diag(d)*A (copy) */
OrCsr *or_amg_smooth_prolongator(const OrCsr *A, const OrCsr *Pt, double omega); /* Pt - omega D^-1 A Pt */
OrCsr *or_amg_galerkin(const OrCsr *A, const OrCsr *P);                        /* P^T (A P) */
/* bilinear interpolation coarse(Mc x Nc) -> fine(2Mc-1 x 2Nc-1), dof-interleaved; bc!=0 zeroes boundary rows/cols */
OrCsr *or_interp_q1(int Mc, int Nc, int dof, int bc);

/* ---- linear operators / preconditioners ---- */
typedef struct OrOp OrOp;
struct OrOp {
  void (*apply)(OrOp *self, const double *x, double *y);
  void (*destroy)(OrOp *self);
  int n_in, n_out;
  void *ctx;
};
void or_op_apply(OrOp *op, const double *x, double *y);
void or_op_free(OrOp *op);
OrOp *or_op_csr(const OrCsr *A);                         /* y = A x */
OrOp *or_op_jacobi(const OrCsr *A);                      /* y = x ./ diag(A) (0 -> 1), PCJACOBI */
OrOp *or_op_diag_inverse(int n, const double *d);        /* y = x ./ d */
OrOp *or_op_nest(const OrCsr *A00, const OrCsr *A01, const OrCsr *A10, const OrCsr *A11 /* may be NULL */);
OrOp *or_op_schur(const OrCsr *A11 /* may be NULL */, const OrCsr *A10, OrOp *K0, const OrCsr *A01);
/* PCFIELDSPLIT Schur: fact 0 diag,1 lower,2 upper,3 full ; scale applies to DIAG only (PETSc default -1) */
OrOp *or_op_fieldsplit(int fact, const OrCsr *A01, const OrCsr *A10, OrOp *K0, OrOp *KS, double scale);
/* PCLSC: y = Linv (A10 A00 A01) Linv x   [scale_diag: A10 D^-1 A00 D^-1 A01 with D=diag(A00)] */
OrOp *or_op_lsc(const OrCsr *A00, const OrCsr *A01, const OrCsr *A10, OrOp *Linv, int scale_diag);
OrOp *or_op_permuted(OrOp *inner, int n, const int *map); /* y[map[i]] = inner(x[map[.]])[i]: strided fieldsplit */
OrOp *or_op_dense_lu(const OrCsr *A);                    /* exact solve, coarse grid */
/* PCMG V-cycle: nlev levels, level 0 finest.  A[l] operators, P[l] (l=0..nlev-2) coarse(l+1)->fine(l),
 * pre/post smoothers S[l] (KSP-as-op with nonzero-guess support, see or_ksp_as_smoother), coarse solver. */
typedef struct OrKsp OrKsp;
OrOp *or_op_mg(int nlev, const OrCsr **A, const OrCsr **P, OrKsp **smooth, OrOp *coarse);

/* ---- KSP ---- */
enum { OR_KSP_PREONLY = 0, OR_KSP_RICHARDSON = 1, OR_KSP_CHEBYSHEV = 2, OR_KSP_GMRES = 3, OR_KSP_FGMRES = 4, OR_KSP_MINRES = 5 };
enum { OR_CONVERGED_RTOL = 2, OR_CONVERGED_ATOL = 3, OR_CONVERGED_ITS = 4, OR_DIVERGED_ITS = -3, OR_DIVERGED_DTOL = -4,
       OR_DIVERGED_BREAKDOWN = -5, OR_DIVERGED_NANORINF = -9, OR_DIVERGED_INDEFINITE_PC = -8 };
struct OrKsp {
  int type;
  OrOp *A, *M; /* M == NULL: identity */
  double rtol, atol, dtol;
  int max_it, restart;
  int norm_none;        /* 1: no convergence test, run exactly max_it iterations (smoother use) */
  double emin, emax;    /* chebyshev bounds */
  double richardson_scale;
  int orthog;           /* gmres: 0 classical GS (PETSc default), 1 CGS refine_ifneeded, 2 CGS refine_always, 3 modified GS */
  /* results */
  int its, reason;
  double rnorm, rnorm0;
  double *hist; int hist_cap, hist_len; /* residual history (monitor) */
};
OrKsp *or_ksp_create(int type, OrOp *A, OrOp *M);
void or_ksp_free(OrKsp *k);
void or_ksp_set_history(OrKsp *k, int cap);
/* solve A x = b; guess_nonzero=0 zeroes x first (PETSc default) */
int or_ksp_solve(OrKsp *k, const double *b, double *x, int guess_nonzero);
OrOp *or_op_from_ksp(OrKsp *k); /* y = ksp(x), zero initial guess; does not own k */
/* deterministic lambda_max estimate of M^-1 A by 10 power iterations from a hashed start vector;
 * chebyshev bounds = (0.1*est, 1.1*est) as PETSc's default transform */
double or_estimate_lambda_max(OrOp *A, OrOp *M, int nits);
void or_hash_vector(int n, double *v); /* v_i in [0.5,1.5), exact integer hash */

/* ---- vector helpers used by timing legs ---- */
double or_dot(int n, const double *x, const double *y);
double or_norm2(int n, const double *x);
void or_set_threads(int nthreads);
int or_get_threads(void);

#ifdef __cplusplus
}
#endif
#endif
