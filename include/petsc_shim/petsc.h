/*
 * petsc.h -- minimal PETSc API shim (OUR code, not PETSc's): exactly the ~45 PETSc / MPI symbols the
 * reference uses, so that /root/reference/src/{main,SaddlePointProblem,Discretization,
 * Visulaization}.c compile UNCHANGED with  -I include/petsc_shim -I /root/reference/include  and drive
 * libb200sp through its C ABI (include/b200sp.h).  PETSc itself is absent from this image; on a machine that
 * has it, csrc/petsc_plugin.c registers the same back end as real MatType/PCType/KSPType instead.
 *
 * Scope: one process (MPI size 1), 2-D DMDA, AIJ matrices, the KSP call sequence of
 * src/SaddlePointProblem.c:65-72.  Every function returns PetscErrorCode (0 = success) like PETSc.
 */
#ifndef B200SP_PETSC_SHIM_H
#define B200SP_PETSC_SHIM_H

#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- basic types ---- */
typedef int PetscInt;
typedef double PetscScalar;
typedef double PetscReal;
typedef int PetscErrorCode;
typedef int PetscMPIInt;
typedef enum { PETSC_FALSE, PETSC_TRUE } PetscBool;
typedef struct _p_PetscObject *PetscObject;
typedef struct _p_DM *DM;
typedef struct _p_Vec *Vec;
typedef struct _p_Mat *Mat;
typedef struct _p_KSP *KSP;
typedef struct _p_PetscViewer *PetscViewer;
typedef struct _p_ISLocalToGlobalMapping *ISLocalToGlobalMapping;
typedef const char *MatType;
typedef const char *VecType;

typedef enum { NOT_SET_VALUES, INSERT_VALUES, ADD_VALUES } InsertMode;
typedef enum { MAT_FLUSH_ASSEMBLY = 1, MAT_FINAL_ASSEMBLY = 0 } MatAssemblyType;
typedef enum { DM_BOUNDARY_NONE, DM_BOUNDARY_GHOSTED, DM_BOUNDARY_MIRROR, DM_BOUNDARY_PERIODIC } DMBoundaryType;
typedef enum { DMDA_STENCIL_STAR, DMDA_STENCIL_BOX } DMDAStencilType;
typedef enum { DMDA_ELEMENT_P1, DMDA_ELEMENT_Q1 } DMDAElementType;

typedef struct { PetscScalar x, y; } DMDACoor2d;
typedef struct { PetscInt k, j, i, c; } MatStencil;
typedef struct {
  PetscInt dim, dof, sw;
  PetscInt mx, my, mz;
  PetscInt xs, ys, zs;
  PetscInt xm, ym, zm;
  PetscInt gxs, gys, gzs;
  PetscInt gxm, gym, gzm;
  DMBoundaryType bx, by, bz;
  DMDAStencilType st;
  DM da;
} DMDALocalInfo;

#define PETSC_DECIDE (-1)
#define PETSC_DETERMINE (-1)
#define PETSC_DEFAULT (-2)
#define MATAIJ "aij"
#define MATB200SP "b200sp"

#define CHKERRQ(ierr) do { if (ierr) return (ierr); } while (0)
#define PetscMalloc1(n, p) ((*(p) = malloc(((size_t)(n) ? (size_t)(n) : 1) * sizeof(**(p)))) ? 0 : 55)
#define PetscFree(p) (free(p), (p) = NULL, 0)
#define PetscMemzero(p, n) (memset((p), 0, (n)), 0)

/* ---- MPI subset (single process) used by src/Visulaization.c ---- */
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Info;
typedef int MPI_Request;
typedef long long MPI_Offset;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;
typedef struct _p_ShimMPIFile *MPI_File;
#define MPI_COMM_WORLD 0
#define MPI_CHAR 1
#define MPI_INT 2
#define MPI_DOUBLE 3
#define MPI_SUM 1
#define MPI_INFO_NULL 0
#define MPI_MODE_CREATE 1
#define MPI_MODE_WRONLY 4
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_SUCCESS 0
#define PETSC_COMM_WORLD MPI_COMM_WORLD
#define PETSC_COMM_SELF MPI_COMM_WORLD
int MPI_Comm_rank(MPI_Comm, int *);
int MPI_Comm_size(MPI_Comm, int *);
int MPI_File_open(MPI_Comm, const char *, int, MPI_Info, MPI_File *);
int MPI_File_close(MPI_File *);
int MPI_File_write_at(MPI_File, MPI_Offset, const void *, int, MPI_Datatype, MPI_Status *);
int MPI_Ibcast(void *, int, MPI_Datatype, int, MPI_Comm, MPI_Request *);
int MPI_Allreduce(const void *, void *, int, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Isend(const void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *);
int MPI_Send(const void *, int, MPI_Datatype, int, int, MPI_Comm);
int MPI_Recv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status *);

/* ---- Sys ---- */
PetscErrorCode PetscInitialize(int *argc, char ***argv, const char file[], const char help[]);
PetscErrorCode PetscFinalize(void);
PetscErrorCode PetscObjectGetComm(PetscObject obj, MPI_Comm *comm);
PetscErrorCode PetscOptionsSetValue(void *options, const char name[], const char value[]);

/* ---- DM / DMDA ---- */
PetscErrorCode DMDACreate2d(MPI_Comm comm, DMBoundaryType bx, DMBoundaryType by, DMDAStencilType st, PetscInt M, PetscInt N, PetscInt m,
                            PetscInt n, PetscInt dof, PetscInt s, const PetscInt lx[], const PetscInt ly[], DM *da);
PetscErrorCode DMSetMatType(DM dm, MatType t);
PetscErrorCode DMSetFromOptions(DM dm);
PetscErrorCode DMSetUp(DM dm);
PetscErrorCode DMDestroy(DM *dm);
PetscErrorCode DMDASetFieldName(DM da, PetscInt nf, const char name[]);
PetscErrorCode DMDASetUniformCoordinates(DM da, PetscReal xmin, PetscReal xmax, PetscReal ymin, PetscReal ymax, PetscReal zmin, PetscReal zmax);
PetscErrorCode DMGetCoordinateDM(DM dm, DM *cdm);
PetscErrorCode DMGetCoordinatesLocal(DM dm, Vec *c);
PetscErrorCode DMDAVecGetArray(DM da, Vec v, void *array);
PetscErrorCode DMDAVecRestoreArray(DM da, Vec v, void *array);
PetscErrorCode DMDAVecGetArrayRead(DM da, Vec v, void *array);
PetscErrorCode DMDAVecRestoreArrayRead(DM da, Vec v, void *array);
PetscErrorCode DMDAGetElementsCorners(DM da, PetscInt *gx, PetscInt *gy, PetscInt *gz);
PetscErrorCode DMDAGetElementsSizes(DM da, PetscInt *mx, PetscInt *my, PetscInt *mz);
PetscErrorCode DMDAGetLocalInfo(DM da, DMDALocalInfo *info);
PetscErrorCode DMDAGetElements(DM da, PetscInt *nel, PetscInt *nen, const PetscInt *e[]);
PetscErrorCode DMDARestoreElements(DM da, PetscInt *nel, PetscInt *nen, const PetscInt *e[]);
PetscErrorCode DMGetLocalVector(DM dm, Vec *v);
PetscErrorCode DMRestoreLocalVector(DM dm, Vec *v);
PetscErrorCode DMLocalToGlobalBegin(DM dm, Vec l, InsertMode mode, Vec g);
PetscErrorCode DMLocalToGlobalEnd(DM dm, Vec l, InsertMode mode, Vec g);
PetscErrorCode DMGlobalToLocalBegin(DM dm, Vec g, InsertMode mode, Vec l);
PetscErrorCode DMGlobalToLocalEnd(DM dm, Vec g, InsertMode mode, Vec l);
PetscErrorCode DMCreateGlobalVector(DM dm, Vec *v);
PetscErrorCode DMCreateMatrix(DM dm, Mat *A);

/* ---- Mat ---- */
PetscErrorCode MatSetValuesStencil(Mat A, PetscInt m, const MatStencil idxm[], PetscInt n, const MatStencil idxn[], const PetscScalar v[], InsertMode mode);
PetscErrorCode MatSetValues(Mat A, PetscInt m, const PetscInt idxm[], PetscInt n, const PetscInt idxn[], const PetscScalar v[], InsertMode mode);
PetscErrorCode MatAssemblyBegin(Mat A, MatAssemblyType t);
PetscErrorCode MatAssemblyEnd(Mat A, MatAssemblyType t);
PetscErrorCode MatZeroRowsColumns(Mat A, PetscInt n, const PetscInt rows[], PetscScalar diag, Vec x, Vec b);
PetscErrorCode MatViewFromOptions(Mat A, PetscObject obj, const char name[]);
PetscErrorCode MatGetSize(Mat A, PetscInt *m, PetscInt *n);
PetscErrorCode MatDestroy(Mat *A);
/* shim extension (tests / MatView): copy of the assembled CSR; pass NULL arrays to query sizes */
PetscErrorCode MatShimGetCSR(Mat A, PetscInt *nrows, PetscInt *nnz, PetscInt *rowptr, PetscInt *col, PetscScalar *val);

/* ---- Vec ---- */
PetscErrorCode VecZeroEntries(Vec v);
PetscErrorCode VecSet(Vec v, PetscScalar a);
PetscErrorCode VecSetValues(Vec v, PetscInt n, const PetscInt ix[], const PetscScalar y[], InsertMode mode);
PetscErrorCode VecAssemblyBegin(Vec v);
PetscErrorCode VecAssemblyEnd(Vec v);
PetscErrorCode VecViewFromOptions(Vec v, PetscObject obj, const char name[]);
PetscErrorCode VecGetSize(Vec v, PetscInt *n);
PetscErrorCode VecGetArray(Vec v, PetscScalar **a);
PetscErrorCode VecRestoreArray(Vec v, PetscScalar **a);
PetscErrorCode VecDestroy(Vec *v);

/* ---- KSP ---- */
PetscErrorCode KSPCreate(MPI_Comm comm, KSP *ksp);
PetscErrorCode KSPSetOperators(KSP ksp, Mat A, Mat P);
PetscErrorCode KSPSetFromOptions(KSP ksp);
PetscErrorCode KSPSetUp(KSP ksp);
PetscErrorCode KSPSolve(KSP ksp, Vec b, Vec x);
PetscErrorCode KSPGetIterationNumber(KSP ksp, PetscInt *its);
PetscErrorCode KSPGetConvergedReason(KSP ksp, PetscInt *reason);
PetscErrorCode KSPGetResidualNorm(KSP ksp, PetscReal *rnorm);
PetscErrorCode KSPDestroy(KSP *ksp);

#ifdef __cplusplus
}
#endif
#endif
