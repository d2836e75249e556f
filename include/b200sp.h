/*
 * b200sp.h -- C ABI of libb200sp: the B200-native (sm_100a) replacement for the PETSc objects the
 * reference's hot path runs in (Mat / Vec / PC / KSP / DMDA assembly).
 *
 * Every entry point cites the reference call site (relative to /root/reference) or the PETSc routine
 * reached from it that the entry replaces.  Conventions:
 *   - plain C: opaque handles, pointers and sizes; no C++/torch types;
 *   - every function returns int, 0 = success (PetscErrorCode convention, CHKERRQ-compatible);
 *     b200sp_last_error() gives the message of the last failure on the calling thread;
 *   - "host" pointers are ordinary host memory, "dev" pointers are device memory of the context's GPU;
 *   - one host thread per context; calls are synchronous unless documented otherwise;
 *   - there is NO CPU fallback: compute entry points fail with B200SP_ERR_NO_DEVICE without a GPU.
 *     The functions marked [host-only] are pure index arithmetic and work without a GPU.
 */
#ifndef B200SP_H
#define B200SP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200sp_ctx_s *b200sp_ctx;
typedef struct b200sp_vec_s *b200sp_vec;
typedef struct b200sp_mat_s *b200sp_mat;
typedef struct b200sp_ksp_s *b200sp_ksp;
typedef struct b200sp_dmda_s *b200sp_dmda;
typedef struct b200sp_dmda3d_s *b200sp_dmda3d;
typedef struct b200sp_pc_s *b200sp_pc;

enum {
  B200SP_OK = 0,
  B200SP_ERR_ARG = 1,
  B200SP_ERR_NO_DEVICE = 2,
  B200SP_ERR_CUDA = 3,
  B200SP_ERR_UNSUPPORTED = 4,
  B200SP_ERR_NCCL = 5,
  B200SP_ERR_MEM = 6
};

/* KSPConvergedReason values (PETSc numbering) returned by b200sp_ksp_get_converged_reason */
enum {
  B200SP_CONVERGED_RTOL = 2,
  B200SP_CONVERGED_ATOL = 3,
  B200SP_CONVERGED_ITS = 4,
  B200SP_DIVERGED_ITS = -3,
  B200SP_DIVERGED_DTOL = -4,
  B200SP_DIVERGED_BREAKDOWN = -5,
  B200SP_DIVERGED_INDEFINITE_PC = -8,
  B200SP_DIVERGED_NANORINF = -9
};

const char *b200sp_last_error(void);
const char *b200sp_version(void);

/* ---- context: replaces PetscInitialize/PetscFinalize + PETSC_COMM_WORLD (src/main.c:12,17) ----
 * One context per process per GPU.  size>1 joins an NCCL communicator: rank 0 calls
 * b200sp_nccl_unique_id, ships the 128 bytes to the other ranks (any bootstrap: torch.distributed
 * store, file, env), every rank calls b200sp_ctx_create with the same id. */
int b200sp_nccl_unique_id(char id[128]);
int b200sp_ctx_create(int device, int rank, int size, const char nccl_id[128], b200sp_ctx *ctx);
/* In-process rank group: all `size` ranks are THREADS of this process (each with its own context and stream, on
 * one or several GPUs).  Collectives are host barriers + peer copies, so ranks never wait on each other inside
 * a kernel.  Used by the tests to run the whole distributed algorithm on a 1-GPU box, and usable as a
 * single-process multi-GPU mode.  Every collective call must be made by all rank-threads. */
typedef struct b200sp_group_s *b200sp_group;
int b200sp_local_group_create(int size, b200sp_group *group);
int b200sp_local_group_destroy(b200sp_group group);
int b200sp_ctx_create_local(b200sp_group group, int rank, int device, b200sp_ctx *ctx);
int b200sp_ctx_destroy(b200sp_ctx ctx);
int b200sp_ctx_synchronize(b200sp_ctx ctx);
int b200sp_ctx_get_stream(b200sp_ctx ctx, void **cuda_stream);
/* number of kernels of THIS library launched since context creation (bench.py's gpu_launches) */
int b200sp_ctx_get_launch_count(b200sp_ctx ctx, int64_t *count);
/* CUDA-event timing on the context's compute stream (bench.py) */
int b200sp_ctx_timer_start(b200sp_ctx ctx);
int b200sp_ctx_timer_stop(b200sp_ctx ctx, double *milliseconds);
/* per-kernel-class accumulated device time; enable -> every launch is bracketed by events (slow path,
 * measurement only).  classes: "spmv", "mdot", "maxpy", "vec", "assembly", ... ; report as JSON text */
int b200sp_ctx_profile_enable(b200sp_ctx ctx, int on);
int b200sp_ctx_profile_report(b200sp_ctx ctx, char *buf, int buflen);

/* ---- DMDA: replaces DMDACreate2d / DMSetUp / DMDAGetLocalInfo / DMDAGetElementsCorners
 *      (src/Discretization.c:17-25, 144-145, 242) -- 2-D box stencil width 1, no periodicity ---- */
/* [host-only] PETSC_DECIDE process grid and ownership ranges */
int b200sp_dmda_proc_grid(int M, int N, int size, int *m, int *n);
int b200sp_dmda_ownership(int M, int m, int *lx);
/* [host-only] corners of rank's owned node box and its owned-element box */
int b200sp_dmda_corners(int M, int N, int size, int rank, int *xs, int *ys, int *xm, int *ym);
int b200sp_dmda_element_corners(int M, int N, int size, int rank, int *si, int *sj, int *ni, int *nj);
/* [host-only] PETSc global node id of natural node (i,j) in a size-rank layout, and its owner */
int b200sp_dmda_global_node(int M, int N, int size, int i, int j, int *gnode, int *owner);
/* [host-only] halo plan of a rank: ghost node list (PETSc global node ids, ascending = MPIAIJ garray
 * order) with owner ranks, and the per-neighbour send lists (local node ids).  Call with NULL arrays
 * to query counts. */
int b200sp_dmda_halo_plan(int M, int N, int size, int rank, int *nghost, int *ghost_gnode, int *ghost_owner,
                          int *nsend_total, int *send_rank, int *send_lnode);
/* [host-only] the node-keyed view of the same send lists, used when the kernel that PRODUCES a vector pushes its
 * boundary values to the neighbours itself: node_ent[v] = (first entry << 3) | count (<= 7) for owned local node v (0 = not
 * sent, count <= 3); entry e goes to rank entry_rank[e] at position entry_pos[e] of that neighbour's receive range
 * (entry 0 is unused).  Call with NULL arrays to query n_owned / n_entries. */
int b200sp_dmda_halo_push_table(int M, int N, int size, int rank, int *n_owned, int *node_ent, int *n_entries, int *entry_rank, int *entry_pos);
/* M x N nodes = (nx+1) x (ny+1); uses the context's rank/size */
int b200sp_dmda_create(b200sp_ctx ctx, int M, int N, b200sp_dmda *da);
int b200sp_dmda_destroy(b200sp_dmda da);
int b200sp_dmda_get_info(b200sp_dmda da, int *M, int *N, int *xs, int *ys, int *xm, int *ym);

/* ---- Vec: replaces DMCreateGlobalVector / Vec* (src/SaddlePointProblem.c:16,43; src/Discretization.c:264) ---- */
int b200sp_vec_create(b200sp_ctx ctx, int64_t n_local, b200sp_vec *v);
int b200sp_vec_destroy(b200sp_vec v);
int b200sp_vec_get_size(b200sp_vec v, int64_t *n_local);
int b200sp_vec_set(b200sp_vec v, double alpha);                             /* VecSet / VecZeroEntries */
int b200sp_vec_set_values_host(b200sp_vec v, int64_t n, const int *idx, const double *vals); /* VecSetValues INSERT (local ids) */
int b200sp_vec_copy_from_host(b200sp_vec v, const double *host, int64_t n);
int b200sp_vec_copy_to_host(b200sp_vec v, double *host, int64_t n);         /* VecGetArray + copy */
int b200sp_vec_get_device_ptr(b200sp_vec v, double **dev);
int b200sp_vec_copy(b200sp_vec x, b200sp_vec y);                            /* VecCopy: y = x */
int b200sp_vec_scale(b200sp_vec x, double a);                               /* VecScale */
int b200sp_vec_axpy(b200sp_vec y, double a, b200sp_vec x);                  /* VecAXPY: y += a x */
int b200sp_vec_aypx(b200sp_vec y, double a, b200sp_vec x);                  /* VecAYPX: y = x + a y */
int b200sp_vec_waxpy(b200sp_vec w, double a, b200sp_vec x, b200sp_vec y);   /* VecWAXPY: w = a x + y */
int b200sp_vec_pointwise_mult(b200sp_vec w, b200sp_vec x, b200sp_vec y);    /* VecPointwiseMult */
int b200sp_vec_dot(b200sp_vec x, b200sp_vec y, double *result);             /* VecDot (global) */
int b200sp_vec_norm(b200sp_vec x, double *result);                          /* VecNorm NORM_2 (global) */
int b200sp_vec_mdot(b200sp_vec x, int k, const b200sp_vec *y, double *result); /* VecMDot */
int b200sp_vec_maxpy(b200sp_vec y, int k, const double *a, const b200sp_vec *x); /* VecMAXPY */

/* ---- Mat: replaces DMCreateMatrix / MatSetValuesStencil / MatAssembly / MatZeroRowsColumns / MatMult
 *      (src/SaddlePointProblem.c:42; src/Discretization.c:165-169, 268) ---- */
/* CSR from host arrays (local rows; column ids local to the rank: [0,ncols_local) owned, >= ghosts) */
int b200sp_mat_create_csr(b200sp_ctx ctx, int nrows, int ncols, const int *rowptr, const int *col, const double *val, b200sp_mat *A);
/* COO -> CSR on the device: stable radix sort by (row,col), duplicates summed IN INSERTION ORDER
 * (= MatSetValues(ADD_VALUES) then MatAssemblyEnd on a single rank) */
int b200sp_mat_create_coo(b200sp_ctx ctx, int nrows, int ncols, int64_t ncoo, const int *row, const int *col, const double *val, b200sp_mat *A);
int b200sp_mat_destroy(b200sp_mat A);
/* MatSetDM equivalent: the M x N-node, dof-per-node DMDA a square matrix lives on (lets -pc_type mg build its
 * hierarchy when the matrix came from b200sp_mat_create_coo/_csr instead of b200sp_assemble_stress) */
int b200sp_mat_set_grid(b200sp_mat A, int M, int N, int dof);
int b200sp_mat_get_size(b200sp_mat A, int *nrows, int *ncols, int64_t *nnz);
int b200sp_mat_get_csr_host(b200sp_mat A, int *rowptr, int *col, double *val); /* MatView / parity checks */
/* row-length histogram (bins 0,1,2,3-4,5-8,...,1025-2048,>2048: 14 bins) and the SpMV kernel chosen from it:
 * 3 = TMA-staged thread-per-row (short rows, default), 0 = warp-stream (short rows, no TMA),
 * 1 = lanes-per-row vector (medium rows), 2 = block-per-row (long dense rows) */
int b200sp_mat_get_spmv_plan(b200sp_mat A, int64_t hist[14], int *kernel, int *max_row_nnz);
int b200sp_mat_set_spmv_kernel(b200sp_mat A, int kernel); /* override (tests / sweeps) */
/* storage the TMA SpMV actually streams for this matrix (after the first mult): block size of the block-compressed
 * column index (1 x 1 = plain CSR columns), whether the tile-local pattern/value dictionaries are in use (one byte
 * per nonzero + per-tile dictionaries instead of 8-byte values and 4-byte columns), and the resulting matrix bytes
 * per launch (without x and y) */
int b200sp_mat_get_spmv_format(b200sp_mat A, int *block_r, int *block_c, int *value_dict, int64_t *matrix_bytes);
/* choose what the SpMV may derive from the CSR arrays (benchmarks of the general-matrix paths; results are bit-identical
 * in every format): block_index 0 = plain CSR columns, value_dict 0 = plain 8-byte values */
int b200sp_mat_set_spmv_format(b200sp_mat A, int block_index, int value_dict);
int b200sp_mat_mult(b200sp_mat A, b200sp_vec x, b200sp_vec y);              /* MatMult */
int b200sp_mat_mult_add(b200sp_mat A, b200sp_vec x, b200sp_vec y, b200sp_vec z); /* MatMultAdd: z = y + A x */
int b200sp_mat_residual(b200sp_mat A, b200sp_vec b, b200sp_vec x, b200sp_vec r); /* r = b - A x (fused) */
int b200sp_mat_get_diagonal(b200sp_mat A, b200sp_vec d);                    /* MatGetDiagonal */
/* MatMultTranspose: y = A^T x.  The explicit transpose is built on the device once per value state and cached, so the
 * entries of a column are added by ascending row -- the order of PETSc's sequential scatter loop.  Single rank only. */
int b200sp_mat_mult_transpose(b200sp_mat A, b200sp_vec x, b200sp_vec y);
int b200sp_mat_transpose(b200sp_mat A, b200sp_mat *At);                     /* MatTranspose (explicit) */
int b200sp_mat_matmult(b200sp_mat A, b200sp_mat B, b200sp_mat *C);          /* MatMatMult (device SpGEMM) */
/* C = A diag(d) (MatDiagonalScale with a right vector, out of place) and C = A + s B (MatAXPY on the union pattern, out of
 * place); both work on row-partitioned matrices (ghost columns scaled through the halo; B's ghost set must contain A's).
 * They are the pieces of Sp = A11 - A10 diag(A00)^-1 A01 (selfp) and of the LSC operator. */
int b200sp_mat_scale_columns(b200sp_mat A, b200sp_vec d, b200sp_mat *C);
int b200sp_mat_add_scaled(b200sp_mat A, double s, b200sp_mat B, b200sp_mat *C);
/* MatZeroRowsColumns(A,n,rows,diag,NULL,NULL): local row ids; pattern preserved (Appendix A.4) */
/* Aggregation multigrid set-up (-pc_type gamg; PETSc analogue PCGAMG, -pc_gamg_type agg).
 * aggregate: agg_host[node] = aggregate id of every bs-dof node of the square matrix A, -1 for nodes without
 * neighbours (the identity rows MatZeroRowsColumns leaves, src/Discretization.c:268); *nagg = number of aggregates.
 * prolongator: P = P_t - omega D^-1 A P_t with P_t[(i,c),(agg(i),c)] = sqrt(w_i / W_agg(i))   (omega = 0: P_t itself);
 * node_weight[i] = w_i = number of finest-level nodes behind node i (NULL: ones, i.e. A is the finest level),
 * coarse_weight[a] = W_a = sum of w over aggregate a (NULL: not wanted).
 * order: priority of the independent-set selection, 0 = hashed node number (-pc_gamg_mis_ordering hash, the default),
 * 1 = the node number itself (natural: regular aggregates on lexicographically numbered grids, O(grid side) rounds).
 * The algorithm is defined by oracle/sp_oracle_amg.c; aggregates, weights and P_t are bit-identical to it. */
int b200sp_amg_aggregate(b200sp_mat A, int bs, double theta, int order, int *agg_host, int *nagg);
int b200sp_amg_prolongator(b200sp_mat A, int bs, double theta, int order, double omega, const int *node_weight, int *coarse_weight, b200sp_mat *P);
int b200sp_mat_zero_rows_columns(b200sp_mat A, int n, const int *rows, double diag);
int b200sp_mat_zero_rows(b200sp_mat A, int n, const int *rows, double diag); /* MatZeroRows (diag only if square) */
int b200sp_mat_zero_columns(b200sp_mat A, int n, const int *cols);
/* MATNEST 2x2 [A00 A01; A10 A11] acting on [x0; x1] stored contiguously (A11 may be NULL) */
int b200sp_mat_create_nest(b200sp_mat A00, b200sp_mat A01, b200sp_mat A10, b200sp_mat A11, b200sp_mat *K);

/* measurement hook (bench.py --config sweep): time the fused Gram-Schmidt kernels of the GMRES drivers (VecMDot against a
 * strided basis of k vectors; VecMAXPY fused with VecNorm) over `reps` launches each, CUDA events, milliseconds per launch */
int b200sp_bench_orthogonalization(b200sp_ctx ctx, int64_t n, int k, int reps, double *ms_mdot, double *ms_maxpy);

/* ---- device assembly: replaces AssembleOperator_Laplace / AssembleRHS_Laplace / ApplyBC_Laplace and
 *      the stubbed AssembleOperator_Constraints (src/Discretization.c:130-290) ---- */
/* A: DMCreateMatrix pattern + element stress matrices summed in the reference's element order */
int b200sp_assemble_stress(b200sp_dmda da, int as_written, b200sp_mat *A);
/* the same operator with a coefficient per Gauss point -- an input of FormStressOperatorQ12D (src/Discretization.c:151-157
 * sets it to 1): coeff_kind 0 = 1.0, 1 = the smooth viscosity 1 + x(1-y)/2 (every stored value distinct: the SpMV's
 * dictionaries decline and the block-index kernel runs; used for the variable-coefficient measurements) */
int b200sp_assemble_stress_coeff(b200sp_dmda da, int coeff_kind, b200sp_mat *A);
/* f: rhs_kind 0 = reference body force (1,2); 1 = rotational force (KKT workloads) */
int b200sp_assemble_rhs(b200sp_dmda da, int as_written, int rhs_kind, b200sp_vec f);
/* KKT blocks on the same nodal grid: Bt (gradient), B = Bt^T (divergence), C (stabilisation, the (2,2)
 * block), Q (= -pressure mass matrix, the "user" Schur preconditioning matrix) */
int b200sp_assemble_kkt(b200sp_dmda da, b200sp_mat *Bt, b200sp_mat *B, b200sp_mat *C, b200sp_mat *Q);
/* ---- 3-D (BASELINE config 4; the reference is 2-D only, include/Discretization.h:8): DMDACreate3d-style grid of M x N x P
 *      nodes, box stencil, PETSC_DECIDE process grid, and the Q1-hexahedron analogue of the 2-D assembly (definition:
 *      oracle/sp_oracle3d.c).  Velocity 3 dof per node, pressure 1; rows are local, columns local (ghosts after the owned). ---- */
int b200sp_dmda3d_proc_grid(int M, int N, int P, int size, int *m, int *n, int *p);
int b200sp_dmda3d_corners(int M, int N, int P, int size, int rank, int *xs, int *ys, int *zs, int *xm, int *ym, int *zm); /* host only */
int b200sp_dmda3d_global_node(int M, int N, int P, int size, int i, int j, int k, int *gnode, int *owner);                /* host only */
int b200sp_dmda3d_create(b200sp_ctx ctx, int M, int N, int P, b200sp_dmda3d *da);
int b200sp_dmda3d_destroy(b200sp_dmda3d da);
int b200sp_dmda3d_get_info(b200sp_dmda3d da, int *xs, int *ys, int *zs, int *xm, int *ym, int *zm, int64_t *first_global_node);
int b200sp_dmda3d_bc_ids(b200sp_dmda3d da, int dof, int *n, int *ids); /* local row ids of the owned boundary nodes */
int b200sp_assemble3d_stress(b200sp_dmda3d da, b200sp_mat *A);
int b200sp_assemble3d_rhs(b200sp_dmda3d da, int rhs_kind, b200sp_vec f); /* 0: (1,2,3); 1: rotational force about z */
int b200sp_assemble3d_kkt(b200sp_dmda3d da, b200sp_mat *Bt, b200sp_mat *B, b200sp_mat *C, b200sp_mat *Q);

/* AssembleOperator_Constraints (an empty stub in the reference, src/Discretization.c:277-283; B is 4 x nCols,
 * src/SaddlePointProblem.c:48-49): the four dense constraint rows "barycentre and volume" (src/main.c:1) -- barycentre
 * x / y, dilation moment, rotation moment of the displacement field about the domain centre -- and B^T.  One rank only. */
int b200sp_assemble_constraints(b200sp_dmda da, b200sp_mat *B, b200sp_mat *Bt);
/* ApplyBC_Laplace id list (local row ids of this rank, ascending); ids may be NULL to query the count */
int b200sp_dmda_bc_ids(b200sp_dmda da, int dof, int *n, int *ids);
/* Q1 interpolation coarse->fine (DMCreateInterpolation on a DMDA); bc!=0 zeroes Dirichlet rows/cols */
int b200sp_interp_q1(b200sp_ctx ctx, int Mc, int Nc, int dof, int bc, b200sp_mat *P);

/* ---- KSP (+PC): replaces KSPCreate / KSPSetOperators / KSPSetFromOptions / KSPSetUp / KSPSolve /
 *      KSPDestroy (src/SaddlePointProblem.c:65-72) ---- */
int b200sp_ksp_create(b200sp_ctx ctx, b200sp_ksp *ksp);
int b200sp_ksp_destroy(b200sp_ksp *ksp);                                    /* nulls the handle like KSPDestroy */
int b200sp_ksp_set_operators(b200sp_ksp ksp, b200sp_mat Amat, b200sp_mat Pmat);
/* PETSc options-database text, e.g. "-ksp_type fgmres -ksp_rtol 1e-8 -pc_type fieldsplit ..." (Appendix A.8) */
int b200sp_ksp_set_options(b200sp_ksp ksp, const char *options);
/* PCFieldSplitSetSchurPre(pc, PC_FIELDSPLIT_SCHUR_PRE_USER, Q) */
int b200sp_ksp_set_schur_user_mat(b200sp_ksp ksp, b200sp_mat Q);
/* grid the velocity block lives on (needed by -pc_type mg: KSPSetDM equivalent) */
int b200sp_ksp_set_dmda(b200sp_ksp ksp, b200sp_dmda da);
int b200sp_ksp_setup(b200sp_ksp ksp);
int b200sp_ksp_solve(b200sp_ksp ksp, b200sp_vec b, b200sp_vec x);
/* same solve with HOST buffers: H2D of b and D2H of x inside the call (the e2e path of bench.py) */
int b200sp_ksp_solve_host(b200sp_ksp ksp, const double *b_host, double *x_host, int64_t n);
int b200sp_ksp_get_iteration_number(b200sp_ksp ksp, int *its);
int b200sp_ksp_get_residual_norm(b200sp_ksp ksp, double *rnorm);
int b200sp_ksp_get_converged_reason(b200sp_ksp ksp, int *reason);
int b200sp_ksp_get_residual_history(b200sp_ksp ksp, double *hist, int cap, int *len);
/* ---- PC as an object of its own: what a PETSc PCRegister'ed type binds (PCSetOperators / PCSetFromOptions /
 *      PCSetUp / PCApply / PCView / PCDestroy); the reference reaches it through KSPSetFromOptions
 *      (src/SaddlePointProblem.c:67) and -pc_type.  Same options text, same setup, same kernels as inside a KSP. ---- */
int b200sp_pc_create(b200sp_ctx ctx, b200sp_pc *pc);
int b200sp_pc_destroy(b200sp_pc *pc);
int b200sp_pc_set_operators(b200sp_pc pc, b200sp_mat Amat, b200sp_mat Pmat);
int b200sp_pc_set_options(b200sp_pc pc, const char *options);      /* "-pc_type fieldsplit -pc_fieldsplit_schur_fact_type upper ..." */
int b200sp_pc_set_schur_user_mat(b200sp_pc pc, b200sp_mat Q);
int b200sp_pc_set_dmda(b200sp_pc pc, b200sp_dmda da);
int b200sp_pc_setup(b200sp_pc pc);
int b200sp_pc_apply(b200sp_pc pc, b200sp_vec x, b200sp_vec y);     /* y = M^-1 x */
int b200sp_pc_view(b200sp_pc pc, char *buf, int buflen);
/* apply only the preconditioner / only the operator (parity tests of PCApply / MatMult on nests) */
int b200sp_ksp_pc_apply(b200sp_ksp ksp, b200sp_vec x, b200sp_vec y);
int b200sp_ksp_view(b200sp_ksp ksp, char *buf, int buflen);

#ifdef __cplusplus
}
#endif
#endif
