/* petsc_plugin.c -- real PETSc registration glue for libb200sp.
 *
 * Compiled ONLY where PETSc's headers exist (-DB200SP_HAVE_PETSC, against the private headers of the installed PETSc,
 * 3.12 ... 3.20 API); in this image PETSc is absent, so the translation unit is empty and the file is carried as source.
 *     mpicc -fPIC -shared -DB200SP_HAVE_PETSC $(pkg-config --cflags PETSc) -I include \
 *           saddle_point_petsc_b200/csrc/petsc_plugin.c -L saddle_point_petsc_b200 -lb200sp -o libb200sp_petsc.so
 *     ./saddle_point_run -dll_append ./libb200sp_petsc.so -dm_mat_type b200sp -ksp_type b200sp \
 *           -b200sp_ksp_type fgmres -b200sp_pc_type fieldsplit ...        (or  -pc_type b200sp  under any PETSc KSP)
 * PETSc calls PetscDLLibraryRegister_b200sp() when it loads the library; nothing in the reference changes: the types are
 * selected through options the reference already honours (DMSetFromOptions, src/Discretization.c:20 -> -dm_mat_type;
 * KSPSetFromOptions, src/SaddlePointProblem.c:67 -> -ksp_type / -pc_type).
 *
 * What is registered (one rank; the row-partitioned path needs the DMDA layout):
 *   MATB200SP  "b200sp"  a MATSEQAIJ subclass (the pattern of MATSEQAIJCUSPARSE): assembly stays PETSc's, so
 *                        MatSetValuesStencil / MatAssembly / MatZeroRowsColumns (src/Discretization.c:165-169, 268) keep
 *                        PETSc's exact semantics; MatMult / MatMultTranspose / MatGetDiagonal run on the device copy,
 *                        refreshed from the AIJ arrays whenever the object state changes.
 *   PCB200SP   "b200sp"  PCSetUp -> b200sp_pc_setup, PCApply -> b200sp_pc_apply (fieldsplit-Schur / LSC / MG / Jacobi,
 *                        chosen by the -b200sp_* options), usable under any PETSc KSP.
 *   KSPB200SP  "b200sp"  the whole preconditioned Krylov solve on the device (KSPSolve -> b200sp_ksp_solve_host).
 * Options: every PETSc option spelled -b200sp_<name> is handed to the library as -<name> (e.g. -b200sp_ksp_type fgmres,
 * -b200sp_pc_fieldsplit_schur_fact_type upper, -b200sp_fieldsplit_0_pc_type mg), so the library's own -ksp_* / -pc_*
 * namespace cannot collide with the PETSc objects that host it.
 */
#ifdef B200SP_HAVE_PETSC
#include <petsc/private/matimpl.h>
#include <petsc/private/kspimpl.h>
#include <petsc/private/pcimpl.h>
#include <../src/mat/impls/aij/seq/aij.h>
#include <petscdmda.h>
#include "b200sp.h"

#define MATB200SP "b200sp"
#define PCB200SP "b200sp"
#define KSPB200SP "b200sp"

static b200sp_ctx g_ctx = NULL;

#define B2CHK(call)                                                                                        \
  do {                                                                                                     \
    int rc_ = (call);                                                                                      \
    if (rc_) SETERRQ2(PETSC_COMM_SELF, PETSC_ERR_LIB, "libb200sp error %d: %s", rc_, b200sp_last_error()); \
  } while (0)

static PetscErrorCode B200SPContext(void)
{
  PetscFunctionBegin;
  if (!g_ctx) {
    PetscInt  dev = 0;
    PetscBool set;
    PetscErrorCode ierr = PetscOptionsGetInt(NULL, NULL, "-b200sp_device", &dev, &set);CHKERRQ(ierr);
    B2CHK(b200sp_ctx_create((int)dev, 0, 1, NULL, &g_ctx)); /* fails loudly without a B200: no CPU fallback */
  }
  PetscFunctionReturn(0);
}

/* every option -b200sp_<name> [value] of the PETSc options database, as "-<name> value ..." text for the library */
static PetscErrorCode B200SPOptionsText(char **text)
{
  char          *all, *tok, *out, *save = NULL;
  size_t         len;
  PetscErrorCode ierr;

  PetscFunctionBegin;
  ierr = PetscOptionsGetAll(NULL, &all);CHKERRQ(ierr);
  ierr = PetscStrlen(all, &len);CHKERRQ(ierr);
  ierr = PetscMalloc1(len + 2, &out);CHKERRQ(ierr);
  out[0] = 0;
  for (tok = strtok_r(all, " ", &save); tok; tok = strtok_r(NULL, " ", &save)) {
    if (!strncmp(tok, "-b200sp_", 8) && strcmp(tok, "-b200sp_device")) {
      strcat(out, "-"); strcat(out, tok + 8); strcat(out, " ");
      /* the value, if the next token is not another option name */
      char *peek = save;
      while (peek && *peek == ' ') ++peek;
      if (peek && *peek && !(peek[0] == '-' && !(peek[1] >= '0' && peek[1] <= '9') && peek[1] != '.')) {
        tok = strtok_r(NULL, " ", &save);
        strcat(out, tok); strcat(out, " ");
      }
    }
  }
  ierr = PetscFree(all);CHKERRQ(ierr);
  *text = out;
  PetscFunctionReturn(0);
}

/* ------------------------------------------------------------------ MATB200SP: MATSEQAIJ + a device copy */
typedef struct {
  b200sp_mat       m;     /* device CSR built from the AIJ arrays */
  PetscObjectState state; /* object state the copy was taken at */
  b200sp_vec       x, y;  /* staging vectors for host Vec arrays */
} MatB200SP;

static PetscErrorCode MatB200SPGet(Mat A, MatB200SP **out)
{
  PetscContainer c;
  PetscErrorCode ierr;

  PetscFunctionBegin;
  ierr = PetscObjectQuery((PetscObject)A, "b200sp_data", (PetscObject *)&c);CHKERRQ(ierr);
  if (!c) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_ARG_WRONGSTATE, "matrix is not of type b200sp");
  ierr = PetscContainerGetPointer(c, (void **)out);CHKERRQ(ierr);
  PetscFunctionReturn(0);
}
static PetscErrorCode MatB200SPDestroyData(void *p)
{
  MatB200SP *d = (MatB200SP *)p;
  if (d->m) b200sp_mat_destroy(d->m);
  if (d->x) b200sp_vec_destroy(d->x);
  if (d->y) b200sp_vec_destroy(d->y);
  return PetscFree(d);
}
/* (re)build the device matrix when the AIJ values changed since the last copy (PetscObjectState) */
static PetscErrorCode MatB200SPSync(Mat A, b200sp_mat *m)
{
  MatB200SP       *d;
  Mat_SeqAIJ      *a = (Mat_SeqAIJ *)A->data;
  PetscObjectState st;
  PetscErrorCode   ierr;

  PetscFunctionBegin;
  ierr = MatB200SPGet(A, &d);CHKERRQ(ierr);
  ierr = PetscObjectStateGet((PetscObject)A, &st);CHKERRQ(ierr);
  if (!d->m || st != d->state) {
    DM       dm;
    PetscInt bs;
    if (d->m) { B2CHK(b200sp_mat_destroy(d->m)); d->m = NULL; }
    /* PetscInt must be 32-bit (the reference passes PetscInt* as int*, src/Visulaization.c:17) */
    B2CHK(b200sp_mat_create_csr(g_ctx, (int)A->rmap->n, (int)A->cmap->n, (const int *)a->i, (const int *)a->j, a->a, &d->m));
    ierr = MatGetDM(A, &dm);CHKERRQ(ierr);
    ierr = MatGetBlockSize(A, &bs);CHKERRQ(ierr);
    if (dm) { /* DMCreateMatrix on a DMDA (src/SaddlePointProblem.c:42): tell the library the grid, for -pc_type mg / fieldsplit */
      PetscInt  dim, M, N, dof;
      PetscBool isda;
      ierr = PetscObjectTypeCompare((PetscObject)dm, DMDA, &isda);CHKERRQ(ierr);
      if (isda) {
        ierr = DMDAGetInfo(dm, &dim, &M, &N, NULL, NULL, NULL, NULL, &dof, NULL, NULL, NULL, NULL, NULL);CHKERRQ(ierr);
        if (dim == 2 && M * N * dof == A->rmap->n) B2CHK(b200sp_mat_set_grid(d->m, (int)M, (int)N, (int)dof));
      }
    }
    d->state = st;
  }
  *m = d->m;
  PetscFunctionReturn(0);
}
static PetscErrorCode B200SPStage(b200sp_vec *v, PetscInt n)
{
  int64_t have = -1;
  PetscFunctionBegin;
  if (*v) b200sp_vec_get_size(*v, &have);
  if (have != (int64_t)n) {
    if (*v) B2CHK(b200sp_vec_destroy(*v));
    B2CHK(b200sp_vec_create(g_ctx, (int64_t)n, v));
  }
  PetscFunctionReturn(0);
}
static PetscErrorCode MatMultKernel_B200SP(Mat A, Vec x, Vec y, PetscBool transpose)
{
  MatB200SP         *d;
  b200sp_mat         m;
  const PetscScalar *xa;
  PetscScalar       *ya;
  PetscInt           nx, ny;
  PetscErrorCode     ierr;

  PetscFunctionBegin;
  ierr = MatB200SPSync(A, &m);CHKERRQ(ierr);
  ierr = MatB200SPGet(A, &d);CHKERRQ(ierr);
  ierr = VecGetLocalSize(x, &nx);CHKERRQ(ierr);
  ierr = VecGetLocalSize(y, &ny);CHKERRQ(ierr);
  ierr = B200SPStage(&d->x, nx);CHKERRQ(ierr);
  ierr = B200SPStage(&d->y, ny);CHKERRQ(ierr);
  ierr = VecGetArrayRead(x, &xa);CHKERRQ(ierr);
  B2CHK(b200sp_vec_copy_from_host(d->x, xa, (int64_t)nx));
  ierr = VecRestoreArrayRead(x, &xa);CHKERRQ(ierr);
  if (transpose) B2CHK(b200sp_mat_mult_transpose(m, d->x, d->y));
  else B2CHK(b200sp_mat_mult(m, d->x, d->y));
  ierr = VecGetArray(y, &ya);CHKERRQ(ierr);
  B2CHK(b200sp_vec_copy_to_host(d->y, ya, (int64_t)ny));
  ierr = VecRestoreArray(y, &ya);CHKERRQ(ierr);
  PetscFunctionReturn(0);
}
static PetscErrorCode MatMult_B200SP(Mat A, Vec x, Vec y) { return MatMultKernel_B200SP(A, x, y, PETSC_FALSE); }
static PetscErrorCode MatMultTranspose_B200SP(Mat A, Vec x, Vec y) { return MatMultKernel_B200SP(A, x, y, PETSC_TRUE); }

PETSC_EXTERN PetscErrorCode MatCreate_B200SP(Mat A)
{
  MatB200SP     *d;
  PetscContainer c;
  PetscErrorCode ierr;

  PetscFunctionBegin;
  ierr = B200SPContext();CHKERRQ(ierr);
  ierr = MatSetType(A, MATSEQAIJ);CHKERRQ(ierr); /* storage, MatSetValues*, assembly, MatZeroRowsColumns: PETSc's own */
  ierr = PetscNew(&d);CHKERRQ(ierr);
  ierr = PetscContainerCreate(PETSC_COMM_SELF, &c);CHKERRQ(ierr);
  ierr = PetscContainerSetPointer(c, d);CHKERRQ(ierr);
  ierr = PetscContainerSetUserDestroy(c, MatB200SPDestroyData);CHKERRQ(ierr);
  ierr = PetscObjectCompose((PetscObject)A, "b200sp_data", (PetscObject)c);CHKERRQ(ierr);
  ierr = PetscContainerDestroy(&c);CHKERRQ(ierr);
  A->ops->mult          = MatMult_B200SP;
  A->ops->multtranspose = MatMultTranspose_B200SP;
  ierr = PetscObjectChangeTypeName((PetscObject)A, MATB200SP);CHKERRQ(ierr);
  PetscFunctionReturn(0);
}

/* device handle of an operator: a MATB200SP gives its copy; any other assembled SeqAIJ matrix is copied once per state */
static PetscErrorCode B200SPOperator(Mat A, b200sp_mat *m, b200sp_mat *owned)
{
  PetscBool      ours, aij;
  PetscErrorCode ierr;

  PetscFunctionBegin;
  *owned = NULL;
  ierr = PetscObjectTypeCompare((PetscObject)A, MATB200SP, &ours);CHKERRQ(ierr);
  if (ours) { ierr = MatB200SPSync(A, m);CHKERRQ(ierr); PetscFunctionReturn(0); }
  ierr = PetscObjectTypeCompare((PetscObject)A, MATSEQAIJ, &aij);CHKERRQ(ierr);
  if (!aij) SETERRQ(PetscObjectComm((PetscObject)A), PETSC_ERR_SUP, "b200sp needs a b200sp or seqaij operator (use -dm_mat_type b200sp)");
  {
    Mat_SeqAIJ *a = (Mat_SeqAIJ *)A->data;
    B2CHK(b200sp_mat_create_csr(g_ctx, (int)A->rmap->n, (int)A->cmap->n, (const int *)a->i, (const int *)a->j, a->a, owned));
    *m = *owned;
  }
  PetscFunctionReturn(0);
}

/* ------------------------------------------------------------------ PCB200SP */
typedef struct { b200sp_pc pc; b200sp_mat ownA, ownP; b200sp_vec x, y; } PCB200SPData;

static PetscErrorCode PCSetUp_B200SP(PC pc)
{
  PCB200SPData  *d = (PCB200SPData *)pc->data;
  Mat            A, P;
  b200sp_mat     mA, mP;
  char          *opts;
  PetscErrorCode ierr;

  PetscFunctionBegin;
  ierr = PCGetOperators(pc, &A, &P);CHKERRQ(ierr);
  if (d->ownA) { b200sp_mat_destroy(d->ownA); d->ownA = NULL; }
  if (d->ownP) { b200sp_mat_destroy(d->ownP); d->ownP = NULL; }
  ierr = B200SPOperator(A, &mA, &d->ownA);CHKERRQ(ierr);
  if (P == A) mP = mA; else { ierr = B200SPOperator(P, &mP, &d->ownP);CHKERRQ(ierr); }
  ierr = B200SPOptionsText(&opts);CHKERRQ(ierr);
  B2CHK(b200sp_pc_set_operators(d->pc, mA, mP));
  B2CHK(b200sp_pc_set_options(d->pc, opts));
  ierr = PetscFree(opts);CHKERRQ(ierr);
  B2CHK(b200sp_pc_setup(d->pc));
  PetscFunctionReturn(0);
}
static PetscErrorCode PCApply_B200SP(PC pc, Vec x, Vec y)
{
  PCB200SPData      *d = (PCB200SPData *)pc->data;
  const PetscScalar *xa;
  PetscScalar       *ya;
  PetscInt           n;
  PetscErrorCode     ierr;

  PetscFunctionBegin;
  ierr = VecGetLocalSize(x, &n);CHKERRQ(ierr);
  ierr = B200SPStage(&d->x, n);CHKERRQ(ierr);
  ierr = B200SPStage(&d->y, n);CHKERRQ(ierr);
  ierr = VecGetArrayRead(x, &xa);CHKERRQ(ierr);
  B2CHK(b200sp_vec_copy_from_host(d->x, xa, (int64_t)n));
  ierr = VecRestoreArrayRead(x, &xa);CHKERRQ(ierr);
  B2CHK(b200sp_pc_apply(d->pc, d->x, d->y));
  ierr = VecGetArray(y, &ya);CHKERRQ(ierr);
  B2CHK(b200sp_vec_copy_to_host(d->y, ya, (int64_t)n));
  ierr = VecRestoreArray(y, &ya);CHKERRQ(ierr);
  PetscFunctionReturn(0);
}
static PetscErrorCode PCView_B200SP(PC pc, PetscViewer viewer)
{
  PCB200SPData  *d = (PCB200SPData *)pc->data;
  char           buf[16384];
  PetscBool      ascii;
  PetscErrorCode ierr;

  PetscFunctionBegin;
  ierr = PetscObjectTypeCompare((PetscObject)viewer, PETSCVIEWERASCII, &ascii);CHKERRQ(ierr);
  if (ascii) { B2CHK(b200sp_pc_view(d->pc, buf, (int)sizeof(buf))); ierr = PetscViewerASCIIPrintf(viewer, "%s", buf);CHKERRQ(ierr); }
  PetscFunctionReturn(0);
}
static PetscErrorCode PCDestroy_B200SP(PC pc)
{
  PCB200SPData *d = (PCB200SPData *)pc->data;
  PetscFunctionBegin;
  if (d->pc) b200sp_pc_destroy(&d->pc);
  if (d->ownA) b200sp_mat_destroy(d->ownA);
  if (d->ownP) b200sp_mat_destroy(d->ownP);
  if (d->x) b200sp_vec_destroy(d->x);
  if (d->y) b200sp_vec_destroy(d->y);
  PetscFunctionReturn(PetscFree(pc->data));
}
PETSC_EXTERN PetscErrorCode PCCreate_B200SP(PC pc)
{
  PCB200SPData  *d;
  PetscErrorCode ierr;

  PetscFunctionBegin;
  ierr = B200SPContext();CHKERRQ(ierr);
  ierr = PetscNew(&d);CHKERRQ(ierr);
  B2CHK(b200sp_pc_create(g_ctx, &d->pc));
  pc->data         = (void *)d;
  pc->ops->setup   = PCSetUp_B200SP;
  pc->ops->apply   = PCApply_B200SP;
  pc->ops->view    = PCView_B200SP;
  pc->ops->destroy = PCDestroy_B200SP;
  PetscFunctionReturn(0);
}

/* ------------------------------------------------------------------ KSPB200SP: the whole solve on the device */
typedef struct { b200sp_ksp ksp; b200sp_mat ownA, ownP; } KSPB200SPData;

static PetscErrorCode KSPSetUp_B200SP(KSP ksp)
{
  KSPB200SPData *d = (KSPB200SPData *)ksp->data;
  Mat            A, P;
  b200sp_mat     mA, mP;
  char          *opts;
  PetscErrorCode ierr;

  PetscFunctionBegin;
  ierr = KSPGetOperators(ksp, &A, &P);CHKERRQ(ierr); /* KSPSetOperators(ksp, A, A), src/SaddlePointProblem.c:66 */
  if (d->ownA) { b200sp_mat_destroy(d->ownA); d->ownA = NULL; }
  if (d->ownP) { b200sp_mat_destroy(d->ownP); d->ownP = NULL; }
  ierr = B200SPOperator(A, &mA, &d->ownA);CHKERRQ(ierr);
  if (P == A) mP = mA; else { ierr = B200SPOperator(P, &mP, &d->ownP);CHKERRQ(ierr); }
  ierr = B200SPOptionsText(&opts);CHKERRQ(ierr);
  B2CHK(b200sp_ksp_set_operators(d->ksp, mA, mP));
  B2CHK(b200sp_ksp_set_options(d->ksp, opts));
  ierr = PetscFree(opts);CHKERRQ(ierr);
  B2CHK(b200sp_ksp_setup(d->ksp));
  PetscFunctionReturn(0);
}
static PetscErrorCode KSPSolve_B200SP(KSP ksp)
{
  KSPB200SPData     *d = (KSPB200SPData *)ksp->data;
  const PetscScalar *b;
  PetscScalar       *x;
  PetscInt           n;
  int                its = 0, reason = 0;
  double             rnorm = 0.0;
  PetscErrorCode     ierr;

  PetscFunctionBegin;
  ierr = VecGetLocalSize(ksp->vec_rhs, &n);CHKERRQ(ierr);
  ierr = VecGetArrayRead(ksp->vec_rhs, &b);CHKERRQ(ierr);
  ierr = VecGetArray(ksp->vec_sol, &x);CHKERRQ(ierr);
  B2CHK(b200sp_ksp_solve_host(d->ksp, b, x, (int64_t)n)); /* zero initial guess, like the reference's KSPSolve (:70) */
  ierr = VecRestoreArray(ksp->vec_sol, &x);CHKERRQ(ierr);
  ierr = VecRestoreArrayRead(ksp->vec_rhs, &b);CHKERRQ(ierr);
  B2CHK(b200sp_ksp_get_iteration_number(d->ksp, &its));
  B2CHK(b200sp_ksp_get_converged_reason(d->ksp, &reason));
  B2CHK(b200sp_ksp_get_residual_norm(d->ksp, &rnorm));
  ksp->its    = its;
  ksp->rnorm  = rnorm;
  ksp->reason = (KSPConvergedReason)reason; /* the library uses PETSc's numeric values (include/b200sp.h) */
  PetscFunctionReturn(0);
}
static PetscErrorCode KSPView_B200SP(KSP ksp, PetscViewer viewer)
{
  KSPB200SPData *d = (KSPB200SPData *)ksp->data;
  char           buf[16384];
  PetscBool      ascii;
  PetscErrorCode ierr;

  PetscFunctionBegin;
  ierr = PetscObjectTypeCompare((PetscObject)viewer, PETSCVIEWERASCII, &ascii);CHKERRQ(ierr);
  if (ascii) { B2CHK(b200sp_ksp_view(d->ksp, buf, (int)sizeof(buf))); ierr = PetscViewerASCIIPrintf(viewer, "%s", buf);CHKERRQ(ierr); }
  PetscFunctionReturn(0);
}
static PetscErrorCode KSPDestroy_B200SP(KSP ksp)
{
  KSPB200SPData *d = (KSPB200SPData *)ksp->data;
  PetscFunctionBegin;
  if (d->ksp) b200sp_ksp_destroy(&d->ksp);
  if (d->ownA) b200sp_mat_destroy(d->ownA);
  if (d->ownP) b200sp_mat_destroy(d->ownP);
  PetscFunctionReturn(KSPDestroyDefault(ksp));
}
PETSC_EXTERN PetscErrorCode KSPCreate_B200SP(KSP ksp)
{
  KSPB200SPData *d;
  PetscErrorCode ierr;

  PetscFunctionBegin;
  ierr = B200SPContext();CHKERRQ(ierr);
  ierr = PetscNew(&d);CHKERRQ(ierr);
  B2CHK(b200sp_ksp_create(g_ctx, &d->ksp));
  ksp->data = (void *)d;
  /* the library preconditions inside the solve: the hosting KSP has no PC work of its own (use -pc_type none) */
  ierr = KSPSetSupportedNorm(ksp, KSP_NORM_PRECONDITIONED, PC_LEFT, 1);CHKERRQ(ierr);
  ierr = KSPSetSupportedNorm(ksp, KSP_NORM_UNPRECONDITIONED, PC_RIGHT, 1);CHKERRQ(ierr);
  ierr = KSPSetSupportedNorm(ksp, KSP_NORM_NONE, PC_LEFT, 1);CHKERRQ(ierr);
  ksp->ops->setup          = KSPSetUp_B200SP;
  ksp->ops->solve          = KSPSolve_B200SP;
  ksp->ops->view           = KSPView_B200SP;
  ksp->ops->destroy        = KSPDestroy_B200SP;
  ksp->ops->buildsolution  = KSPBuildSolutionDefault;
  ksp->ops->buildresidual  = KSPBuildResidualDefault;
  ksp->ops->setfromoptions = NULL;
  PetscFunctionReturn(0);
}

/* ------------------------------------------------------------------ entry point PETSc calls for -dll_append */
PETSC_EXTERN PetscErrorCode PetscDLLibraryRegister_b200sp(void)
{
  PetscErrorCode ierr;

  PetscFunctionBegin;
  ierr = MatRegister(MATB200SP, MatCreate_B200SP);CHKERRQ(ierr);
  ierr = PCRegister(PCB200SP, PCCreate_B200SP);CHKERRQ(ierr);
  ierr = KSPRegister(KSPB200SP, KSPCreate_B200SP);CHKERRQ(ierr);
  PetscFunctionReturn(0);
}
#else
/* PETSc is not available in this build: see the header comment.  (ISO C forbids an empty translation unit.) */
typedef int b200sp_petsc_plugin_not_built;
#endif
