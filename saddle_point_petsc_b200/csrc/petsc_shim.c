/*
 * petsc_shim.c -- implementation of include/petsc_shim/petsc.h: the host-side PETSc objects the reference
 * touches (options database, DMDA index arithmetic, Vec host arrays, the MatSetValues stash), with every
 * numerical operation delegated to the solver back end behind shim_backend.h.
 *
 * PETSc semantics restated here, single process:
 *   - DMDA: natural ordering == PETSc ordering, ghosted box == whole grid;
 *   - DMCreateMatrix: box-stencil x dof pattern inserted as explicit zeros (so the assembled CSR keeps
 *     PETSc's structure even where no element contributes);
 *   - MatSetValuesStencil: stencil -> global index, out-of-grid stencils dropped, values read ROW-major,
 *     ADD_VALUES recorded in call order; MatAssemblyEnd ships the triplets to the back end, which sorts them
 *     and sums duplicates in that order (device radix sort in libb200sp);
 *   - MatZeroRowsColumns keeps the pattern, sets the diagonal; x=b=NULL only (the reference's use).
 */
#include <petsc.h>
#include <ctype.h>
#include "shim_backend.h"

/* ------------------------------------------------------------------ objects */
struct _p_DM {
  int M, N, dof, sw;
  int is_setup;
  char mattype[32];
  DM cdm;        /* coordinate DM (dof 2) */
  Vec coords;    /* local coordinates */
  int *elements; /* DMDAGetElements cache */
  char fieldname[8][32];
};
struct _p_Vec {
  int n;
  double *a;
  DM dm;
  void **rowptrs; /* DMDAVecGetArray row-pointer table */
};
struct _p_Mat {
  int nrows, ncols;
  DM dm;
  long ncoo, cap;
  int *crow, *ccol;
  double *cval;
  shimbk_mat bk;
};
struct _p_KSP {
  Mat A, P;
  int its, reason;
  double rnorm;
};

#define SHIM_ERR(msg) (fprintf(stderr, "[petsc-shim] %s:%d: %s\n", __FILE__, __LINE__, msg), 1)
#define BK(call) do { if ((call) != 0) { fprintf(stderr, "[petsc-shim] back end '%s' error: %s\n", shimbk_name(), shimbk_last_error()); return 76; } } while (0)

/* ------------------------------------------------------------------ options database */
#define MAX_OPTS 256
static struct { char *name, *value; } g_opts[MAX_OPTS];
static int g_nopts = 0;
static int g_initialized = 0;

static int opt_is_number(const char *s) {
  char *end;
  if (!s || !*s) return 0;
  strtod(s, &end);
  return *end == 0;
}
static void opt_set(const char *name, const char *value) {
  for (int i = 0; i < g_nopts; ++i)
    if (!strcmp(g_opts[i].name, name)) {
      free(g_opts[i].value);
      g_opts[i].value = strdup(value ? value : "");
      return;
    }
  if (g_nopts < MAX_OPTS) {
    g_opts[g_nopts].name = strdup(name);
    g_opts[g_nopts].value = strdup(value ? value : "");
    g_nopts++;
  }
}
static const char *opt_get(const char *name) {
  for (int i = 0; i < g_nopts; ++i)
    if (!strcmp(g_opts[i].name, name)) return g_opts[i].value;
  return NULL;
}
static void opt_parse_tokens(int n, char **tok) {
  for (int i = 0; i < n;) {
    if (tok[i][0] == '-' && !opt_is_number(tok[i])) {
      const char *name = tok[i] + 1;
      if (i + 1 < n && !(tok[i + 1][0] == '-' && !opt_is_number(tok[i + 1]))) { opt_set(name, tok[i + 1]); i += 2; }
      else { opt_set(name, ""); i += 1; }
    } else i += 1;
  }
}
static void opt_parse_string(const char *s) {
  if (!s) return;
  char *copy = strdup(s), *save = NULL, *tok[512];
  int n = 0;
  for (char *t = strtok_r(copy, " \t\n", &save); t && n < 512; t = strtok_r(NULL, " \t\n", &save)) tok[n++] = t;
  opt_parse_tokens(n, tok);
  free(copy);
}
PetscErrorCode PetscOptionsSetValue(void *options, const char name[], const char value[]) {
  (void)options;
  opt_set(name[0] == '-' ? name + 1 : name, value);
  return 0;
}
/* every -ksp_* / -pc_* / -fieldsplit_* option, as options-database text for the back end */
static char *solver_options_text(void) {
  size_t cap = 64;
  for (int i = 0; i < g_nopts; ++i) cap += strlen(g_opts[i].name) + strlen(g_opts[i].value) + 4;
  char *out = (char *)malloc(cap);
  out[0] = 0;
  for (int i = 0; i < g_nopts; ++i) {
    const char *n = g_opts[i].name;
    if (!strncmp(n, "ksp_", 4) || !strncmp(n, "pc_", 3) || !strncmp(n, "fieldsplit_", 11) || !strncmp(n, "mg_", 3)) {
      strcat(out, "-"); strcat(out, n); strcat(out, " ");
      if (g_opts[i].value[0]) { strcat(out, g_opts[i].value); strcat(out, " "); }
    }
  }
  return out;
}

PetscErrorCode PetscInitialize(int *argc, char ***argv, const char file[], const char help[]) {
  (void)file;
  if (g_initialized) return 0;
  if (argc && argv && *argc > 1) opt_parse_tokens(*argc - 1, *argv + 1);
  opt_parse_string(getenv("PETSC_OPTIONS"));
  if (opt_get("help") && help) fputs(help, stdout);
  BK(shimbk_init());
  g_initialized = 1;
  return 0;
}
PetscErrorCode PetscFinalize(void) {
  /* the reference leaks A, f, u and the DM: undisposed objects are tolerated */
  if (g_initialized) BK(shimbk_finalize());
  g_initialized = 0;
  for (int i = 0; i < g_nopts; ++i) { free(g_opts[i].name); free(g_opts[i].value); }
  g_nopts = 0;
  return 0;
}
PetscErrorCode PetscObjectGetComm(PetscObject obj, MPI_Comm *comm) { (void)obj; *comm = MPI_COMM_WORLD; return 0; }

/* ------------------------------------------------------------------ MPI (single process) */
struct _p_ShimMPIFile { FILE *f; };
int MPI_Comm_rank(MPI_Comm c, int *r) { (void)c; *r = 0; return 0; }
int MPI_Comm_size(MPI_Comm c, int *s) { (void)c; *s = 1; return 0; }
int MPI_File_open(MPI_Comm c, const char *name, int amode, MPI_Info info, MPI_File *fh) {
  (void)c; (void)amode; (void)info;
  MPI_File h = (MPI_File)malloc(sizeof(*h));
  h->f = fopen(name, "wb");
  if (!h->f) { free(h); return 1; }
  *fh = h;
  return 0;
}
int MPI_File_close(MPI_File *fh) { if (fh && *fh) { fclose((*fh)->f); free(*fh); *fh = NULL; } return 0; }
int MPI_File_write_at(MPI_File fh, MPI_Offset off, const void *buf, int count, MPI_Datatype t, MPI_Status *st) {
  (void)st;
  size_t sz = t == MPI_CHAR ? 1 : t == MPI_INT ? sizeof(int) : sizeof(double);
  if (fseek(fh->f, (long)off, SEEK_SET)) return 1;
  return fwrite(buf, sz, (size_t)count, fh->f) == (size_t)count ? 0 : 1;
}
int MPI_Ibcast(void *b, int n, MPI_Datatype t, int root, MPI_Comm c, MPI_Request *r) { (void)b; (void)n; (void)t; (void)root; (void)c; if (r) *r = 0; return 0; }
int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) {
  (void)op; (void)c;
  memcpy(r, s, (size_t)n * (t == MPI_CHAR ? 1 : t == MPI_INT ? sizeof(int) : sizeof(double)));
  return 0;
}
int MPI_Isend(const void *b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c, MPI_Request *r) { (void)b; (void)n; (void)t; (void)d; (void)tag; (void)c; if (r) *r = 0; return 0; }
int MPI_Send(const void *b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c) { (void)b; (void)n; (void)t; (void)d; (void)tag; (void)c; return 0; }
int MPI_Recv(void *b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Status *st) { (void)b; (void)n; (void)t; (void)s; (void)tag; (void)c; (void)st; return 0; }

/* ------------------------------------------------------------------ Vec */
static Vec vec_new(int n, DM dm) {
  Vec v = (Vec)calloc(1, sizeof(*v));
  v->n = n;
  v->a = (double *)calloc((size_t)(n ? n : 1), sizeof(double));
  v->dm = dm;
  return v;
}
PetscErrorCode VecZeroEntries(Vec v) { memset(v->a, 0, sizeof(double) * (size_t)v->n); return 0; }
PetscErrorCode VecSet(Vec v, PetscScalar a) { for (int i = 0; i < v->n; ++i) v->a[i] = a; return 0; }
PetscErrorCode VecSetValues(Vec v, PetscInt n, const PetscInt ix[], const PetscScalar y[], InsertMode mode) {
  for (int t = 0; t < n; ++t) {
    if (ix[t] < 0) continue; /* negative indices are ignored, as in PETSc */
    if (ix[t] >= v->n) return SHIM_ERR("VecSetValues: index out of range");
    if (mode == ADD_VALUES) v->a[ix[t]] += y[t]; else v->a[ix[t]] = y[t];
  }
  return 0;
}
PetscErrorCode VecAssemblyBegin(Vec v) { (void)v; return 0; }
PetscErrorCode VecAssemblyEnd(Vec v) { (void)v; return 0; }
PetscErrorCode VecGetSize(Vec v, PetscInt *n) { *n = v->n; return 0; }
PetscErrorCode VecGetArray(Vec v, PetscScalar **a) { *a = v->a; return 0; }
PetscErrorCode VecRestoreArray(Vec v, PetscScalar **a) { (void)v; if (a) *a = NULL; return 0; }
PetscErrorCode VecDestroy(Vec *v) {
  if (v && *v) { free((*v)->a); free((*v)->rowptrs); free(*v); *v = NULL; }
  return 0;
}
PetscErrorCode VecViewFromOptions(Vec v, PetscObject obj, const char name[]) {
  (void)obj;
  if (!opt_get(name + 1)) return 0;
  printf("Vec Object: 1 MPI processes\n  type: b200sp-shim\n");
  for (int i = 0; i < v->n; ++i) printf("%.16g\n", v->a[i]);
  return 0;
}

/* ------------------------------------------------------------------ DM / DMDA */
PetscErrorCode DMDACreate2d(MPI_Comm comm, DMBoundaryType bx, DMBoundaryType by, DMDAStencilType st, PetscInt M, PetscInt N, PetscInt m, PetscInt n,
                            PetscInt dof, PetscInt s, const PetscInt lx[], const PetscInt ly[], DM *da) {
  (void)comm; (void)m; (void)n; (void)lx; (void)ly;
  if (bx != DM_BOUNDARY_NONE || by != DM_BOUNDARY_NONE || st != DMDA_STENCIL_BOX || s != 1) return SHIM_ERR("DMDACreate2d: only non-periodic box stencil of width 1");
  DM d = (DM)calloc(1, sizeof(*d));
  d->M = M; d->N = N; d->dof = dof; d->sw = s;
  strcpy(d->mattype, MATAIJ);
  *da = d;
  return 0;
}
PetscErrorCode DMSetMatType(DM dm, MatType t) { strncpy(dm->mattype, t, sizeof(dm->mattype) - 1); return 0; }
PetscErrorCode DMSetFromOptions(DM dm) {
  const char *v;
  if (dm->is_setup) return SHIM_ERR("DMSetFromOptions after DMSetUp");
  if ((v = opt_get("da_grid_x"))) dm->M = atoi(v);
  if ((v = opt_get("da_grid_y"))) dm->N = atoi(v);
  if ((v = opt_get("dm_mat_type"))) DMSetMatType(dm, v);
  return 0;
}
PetscErrorCode DMSetUp(DM dm) {
  if (dm->M < 2 || dm->N < 2) return SHIM_ERR("DMSetUp: grid needs at least 2 x 2 nodes");
  dm->is_setup = 1;
  return 0;
}
PetscErrorCode DMDestroy(DM *dm) {
  if (dm && *dm) {
    if ((*dm)->cdm) { free((*dm)->cdm); }
    VecDestroy(&(*dm)->coords);
    free((*dm)->elements);
    free(*dm);
    *dm = NULL;
  }
  return 0;
}
PetscErrorCode DMDASetFieldName(DM da, PetscInt nf, const char name[]) {
  if (nf < 0 || nf >= 8) return SHIM_ERR("DMDASetFieldName: field out of range");
  strncpy(da->fieldname[nf], name, 31);
  return 0;
}
PetscErrorCode DMDASetUniformCoordinates(DM da, PetscReal xmin, PetscReal xmax, PetscReal ymin, PetscReal ymax, PetscReal zmin, PetscReal zmax) {
  (void)zmin; (void)zmax;
  if (!da->cdm) {
    da->cdm = (DM)calloc(1, sizeof(*da->cdm));
    da->cdm->M = da->M; da->cdm->N = da->N; da->cdm->dof = 2; da->cdm->sw = da->sw; da->cdm->is_setup = 1;
  }
  VecDestroy(&da->coords);
  da->coords = vec_new(2 * da->M * da->N, da->cdm);
  /* DMDASetUniformCoordinates 2-D: hx = (xmax-xmin)/(M-1); x = xmin + hx*i */
  const double hx = (xmax - xmin) / (double)(da->M - 1), hy = (ymax - ymin) / (double)(da->N - 1);
  for (int j = 0; j < da->N; ++j)
    for (int i = 0; i < da->M; ++i) {
      da->coords->a[2 * (j * da->M + i) + 0] = xmin + hx * (double)i;
      da->coords->a[2 * (j * da->M + i) + 1] = ymin + hy * (double)j;
    }
  return 0;
}
PetscErrorCode DMGetCoordinateDM(DM dm, DM *cdm) { if (!dm->cdm) return SHIM_ERR("no coordinates set"); *cdm = dm->cdm; return 0; }
PetscErrorCode DMGetCoordinatesLocal(DM dm, Vec *c) { if (!dm->coords) return SHIM_ERR("no coordinates set"); *c = dm->coords; return 0; }
/* a[j][i] with GLOBAL node indices; element type is a struct of `dof` scalars (Field, DMDACoor2d) */
PetscErrorCode DMDAVecGetArray(DM da, Vec v, void *array) {
  if (v->n != da->dof * da->M * da->N) return SHIM_ERR("DMDAVecGetArray: vector does not belong to this DMDA");
  if (!v->rowptrs) v->rowptrs = (void **)malloc(sizeof(void *) * (size_t)da->N);
  for (int j = 0; j < da->N; ++j) v->rowptrs[j] = (void *)(v->a + (size_t)j * da->M * da->dof);
  *(void ***)array = v->rowptrs; /* gxs = gys = 0 on one process */
  return 0;
}
PetscErrorCode DMDAVecRestoreArray(DM da, Vec v, void *array) { (void)da; (void)v; if (array) *(void ***)array = NULL; return 0; }
PetscErrorCode DMDAVecGetArrayRead(DM da, Vec v, void *array) { return DMDAVecGetArray(da, v, array); }
PetscErrorCode DMDAVecRestoreArrayRead(DM da, Vec v, void *array) { return DMDAVecRestoreArray(da, v, array); }
PetscErrorCode DMDAGetElementsCorners(DM da, PetscInt *gx, PetscInt *gy, PetscInt *gz) { (void)da; if (gx) *gx = 0; if (gy) *gy = 0; if (gz) *gz = 0; return 0; }
PetscErrorCode DMDAGetElementsSizes(DM da, PetscInt *mx, PetscInt *my, PetscInt *mz) {
  if (mx) *mx = da->M - 1;
  if (my) *my = da->N - 1;
  if (mz) *mz = 0;
  return 0;
}
PetscErrorCode DMDAGetLocalInfo(DM da, DMDALocalInfo *info) {
  memset(info, 0, sizeof(*info));
  info->dim = 2; info->dof = da->dof; info->sw = da->sw;
  info->mx = da->M; info->my = da->N; info->mz = 1;
  info->xm = da->M; info->ym = da->N; info->zm = 1;
  info->gxm = da->M; info->gym = da->N; info->gzm = 1;
  info->st = DMDA_STENCIL_BOX;
  info->da = da;
  return 0;
}
PetscErrorCode DMDAGetElements(DM da, PetscInt *nel, PetscInt *nen, const PetscInt *e[]) {
  const int ne = (da->M - 1) * (da->N - 1);
  if (!da->elements) {
    da->elements = (int *)malloc(sizeof(int) * 4 * (size_t)ne);
    int k = 0;
    for (int j = 0; j < da->N - 1; ++j)
      for (int i = 0; i < da->M - 1; ++i) { /* DMDAGetElements_2D, Q1: counter-clockwise from the lower-left node */
        da->elements[k++] = j * da->M + i;
        da->elements[k++] = j * da->M + i + 1;
        da->elements[k++] = (j + 1) * da->M + i + 1;
        da->elements[k++] = (j + 1) * da->M + i;
      }
  }
  *nel = ne; *nen = 4; *e = da->elements;
  return 0;
}
PetscErrorCode DMDARestoreElements(DM da, PetscInt *nel, PetscInt *nen, const PetscInt *e[]) { (void)da; (void)nel; (void)nen; (void)e; return 0; }
PetscErrorCode DMGetLocalVector(DM dm, Vec *v) { *v = vec_new(dm->dof * dm->M * dm->N, dm); return 0; }
PetscErrorCode DMRestoreLocalVector(DM dm, Vec *v) { (void)dm; return VecDestroy(v); }
PetscErrorCode DMLocalToGlobalBegin(DM dm, Vec l, InsertMode mode, Vec g) {
  (void)dm;
  if (l->n != g->n) return SHIM_ERR("DMLocalToGlobal: size mismatch");
  for (int i = 0; i < g->n; ++i) { if (mode == ADD_VALUES) g->a[i] += l->a[i]; else g->a[i] = l->a[i]; }
  return 0;
}
PetscErrorCode DMLocalToGlobalEnd(DM dm, Vec l, InsertMode mode, Vec g) { (void)dm; (void)l; (void)mode; (void)g; return 0; }
PetscErrorCode DMGlobalToLocalBegin(DM dm, Vec g, InsertMode mode, Vec l) { (void)dm; (void)mode; memcpy(l->a, g->a, sizeof(double) * (size_t)g->n); return 0; }
PetscErrorCode DMGlobalToLocalEnd(DM dm, Vec g, InsertMode mode, Vec l) { (void)dm; (void)g; (void)mode; (void)l; return 0; }
PetscErrorCode DMCreateGlobalVector(DM dm, Vec *v) { if (!dm->is_setup) return SHIM_ERR("DM not set up"); *v = vec_new(dm->dof * dm->M * dm->N, dm); return 0; }

/* ------------------------------------------------------------------ Mat */
static int mat_push(Mat A, int r, int c, double v) {
  if (A->ncoo == A->cap) {
    A->cap = A->cap ? A->cap * 2 : 1024;
    A->crow = (int *)realloc(A->crow, sizeof(int) * (size_t)A->cap);
    A->ccol = (int *)realloc(A->ccol, sizeof(int) * (size_t)A->cap);
    A->cval = (double *)realloc(A->cval, sizeof(double) * (size_t)A->cap);
    if (!A->crow || !A->ccol || !A->cval) return 55;
  }
  A->crow[A->ncoo] = r; A->ccol[A->ncoo] = c; A->cval[A->ncoo] = v;
  A->ncoo++;
  return 0;
}
PetscErrorCode DMCreateMatrix(DM dm, Mat *pA) {
  if (!dm->is_setup) return SHIM_ERR("DM not set up");
  Mat A = (Mat)calloc(1, sizeof(*A));
  A->nrows = A->ncols = dm->dof * dm->M * dm->N;
  A->dm = dm;
  /* preallocation: explicit zeros on the box-stencil x dof pattern (MatSetValues of zeros in DMCreateMatrix_DA_2d_MPIAIJ) */
  for (int j = 0; j < dm->N; ++j)
    for (int i = 0; i < dm->M; ++i)
      for (int c = 0; c < dm->dof; ++c)
        for (int jj = (j > 0 ? j - 1 : 0); jj <= (j < dm->N - 1 ? j + 1 : dm->N - 1); ++jj)
          for (int ii = (i > 0 ? i - 1 : 0); ii <= (i < dm->M - 1 ? i + 1 : dm->M - 1); ++ii)
            for (int cc = 0; cc < dm->dof; ++cc)
              if (mat_push(A, (j * dm->M + i) * dm->dof + c, (jj * dm->M + ii) * dm->dof + cc, 0.0)) return 55;
  PetscErrorCode ierr = MatAssemblyEnd(A, MAT_FINAL_ASSEMBLY);
  if (ierr) return ierr;
  *pA = A;
  return 0;
}
PetscErrorCode MatSetValues(Mat A, PetscInt m, const PetscInt idxm[], PetscInt n, const PetscInt idxn[], const PetscScalar v[], InsertMode mode) {
  if (mode != ADD_VALUES) return SHIM_ERR("MatSetValues: the shim supports ADD_VALUES only (the reference's use)");
  for (int a = 0; a < m; ++a) {
    if (idxm[a] < 0) continue;
    if (idxm[a] >= A->nrows) return SHIM_ERR("MatSetValues: row out of range");
    for (int b = 0; b < n; ++b) {
      if (idxn[b] < 0) continue;
      if (idxn[b] >= A->ncols) return SHIM_ERR("MatSetValues: column out of range");
      if (mat_push(A, idxm[a], idxn[b], v[a * n + b])) return 55; /* row-major read of v */
    }
  }
  return 0;
}
PetscErrorCode MatSetValuesStencil(Mat A, PetscInt m, const MatStencil idxm[], PetscInt n, const MatStencil idxn[], const PetscScalar v[], InsertMode mode) {
  DM dm = A->dm;
  if (!dm) return SHIM_ERR("MatSetValuesStencil: matrix has no DMDA");
  if (m > 64 || n > 64) return SHIM_ERR("MatSetValuesStencil: too many stencil entries");
  PetscInt rm[64], cn[64];
  for (int a = 0; a < m; ++a) { /* stencil -> global index; outside the grid -> -1 (dropped) */
    const MatStencil *s = &idxm[a];
    rm[a] = (s->i < 0 || s->i >= dm->M || s->j < 0 || s->j >= dm->N) ? -1 : (s->j * dm->M + s->i) * dm->dof + s->c;
  }
  for (int b = 0; b < n; ++b) {
    const MatStencil *s = &idxn[b];
    cn[b] = (s->i < 0 || s->i >= dm->M || s->j < 0 || s->j >= dm->N) ? -1 : (s->j * dm->M + s->i) * dm->dof + s->c;
  }
  return MatSetValues(A, m, rm, n, cn, v, mode);
}
PetscErrorCode MatAssemblyBegin(Mat A, MatAssemblyType t) { (void)A; (void)t; return 0; }
PetscErrorCode MatAssemblyEnd(Mat A, MatAssemblyType t) {
  (void)t;
  if (A->bk && A->ncoo == 0) return 0;
  if (A->bk) {
    /* values set after an earlier assembly: the assembled entries come first (they were inserted first) */
    int nr; long nnz;
    BK(shimbk_mat_get_csr(A->bk, &nr, &nnz, NULL, NULL, NULL));
    int *rp = (int *)malloc(sizeof(int) * ((size_t)nr + 1)), *cj = (int *)malloc(sizeof(int) * (size_t)(nnz ? nnz : 1));
    double *va = (double *)malloc(sizeof(double) * (size_t)(nnz ? nnz : 1));
    BK(shimbk_mat_get_csr(A->bk, &nr, &nnz, rp, cj, va));
    long total = nnz + A->ncoo;
    int *r2 = (int *)malloc(sizeof(int) * (size_t)total), *c2 = (int *)malloc(sizeof(int) * (size_t)total);
    double *v2 = (double *)malloc(sizeof(double) * (size_t)total);
    long p = 0;
    for (int r = 0; r < nr; ++r)
      for (int k = rp[r]; k < rp[r + 1]; ++k) { r2[p] = r; c2[p] = cj[k]; v2[p] = va[k]; ++p; }
    memcpy(r2 + p, A->crow, sizeof(int) * (size_t)A->ncoo);
    memcpy(c2 + p, A->ccol, sizeof(int) * (size_t)A->ncoo);
    memcpy(v2 + p, A->cval, sizeof(double) * (size_t)A->ncoo);
    free(rp); free(cj); free(va);
    free(A->crow); free(A->ccol); free(A->cval);
    A->crow = r2; A->ccol = c2; A->cval = v2; A->ncoo = A->cap = total;
    BK(shimbk_mat_destroy(A->bk));
    A->bk = NULL;
  }
  BK(shimbk_mat_from_coo(A->nrows, A->ncols, A->ncoo, A->crow, A->ccol, A->cval, &A->bk));
  if (A->dm) BK(shimbk_mat_set_grid(A->bk, A->dm->M, A->dm->N, A->dm->dof));
  free(A->crow); free(A->ccol); free(A->cval);
  A->crow = A->ccol = NULL; A->cval = NULL; A->ncoo = A->cap = 0;
  return 0;
}
PetscErrorCode MatZeroRowsColumns(Mat A, PetscInt n, const PetscInt rows[], PetscScalar diag, Vec x, Vec b) {
  if (x || b) return SHIM_ERR("MatZeroRowsColumns: x and b must be NULL (the reference's use)");
  if (!A->bk || A->ncoo) return SHIM_ERR("MatZeroRowsColumns: matrix not assembled");
  BK(shimbk_mat_zero_rows_columns(A->bk, n, rows, diag));
  return 0;
}
PetscErrorCode MatGetSize(Mat A, PetscInt *m, PetscInt *n) { if (m) *m = A->nrows; if (n) *n = A->ncols; return 0; }
PetscErrorCode MatShimGetCSR(Mat A, PetscInt *nrows, PetscInt *nnz, PetscInt *rowptr, PetscInt *col, PetscScalar *val) {
  if (!A->bk) return SHIM_ERR("MatShimGetCSR: matrix not assembled");
  int nr; long nz;
  BK(shimbk_mat_get_csr(A->bk, &nr, &nz, rowptr, col, val));
  if (nrows) *nrows = nr;
  if (nnz) *nnz = (PetscInt)nz;
  return 0;
}
PetscErrorCode MatViewFromOptions(Mat A, PetscObject obj, const char name[]) {
  (void)obj;
  if (!opt_get(name + 1)) return 0;
  int nr; long nnz;
  BK(shimbk_mat_get_csr(A->bk, &nr, &nnz, NULL, NULL, NULL));
  int *rp = (int *)malloc(sizeof(int) * ((size_t)nr + 1)), *cj = (int *)malloc(sizeof(int) * (size_t)(nnz ? nnz : 1));
  double *va = (double *)malloc(sizeof(double) * (size_t)(nnz ? nnz : 1));
  BK(shimbk_mat_get_csr(A->bk, &nr, &nnz, rp, cj, va));
  printf("Mat Object: 1 MPI processes\n  type: %s (b200sp-shim, back end %s)\n", A->dm ? A->dm->mattype : "aij", shimbk_name());
  for (int r = 0; r < nr; ++r) {
    printf("row %d:", r);
    for (int k = rp[r]; k < rp[r + 1]; ++k) printf(" (%d, %g) ", cj[k], va[k]);
    printf("\n");
  }
  free(rp); free(cj); free(va);
  return 0;
}
PetscErrorCode MatDestroy(Mat *A) {
  if (A && *A) {
    if ((*A)->bk) shimbk_mat_destroy((*A)->bk);
    free((*A)->crow); free((*A)->ccol); free((*A)->cval);
    free(*A);
    *A = NULL;
  }
  return 0;
}

/* ------------------------------------------------------------------ KSP */
PetscErrorCode KSPCreate(MPI_Comm comm, KSP *ksp) { (void)comm; *ksp = (KSP)calloc(1, sizeof(**ksp)); return 0; }
PetscErrorCode KSPSetOperators(KSP ksp, Mat A, Mat P) { ksp->A = A; ksp->P = P; return 0; }
PetscErrorCode KSPSetFromOptions(KSP ksp) { (void)ksp; return 0; } /* options are read at solve time from the database */
PetscErrorCode KSPSetUp(KSP ksp) { if (!ksp->A) return SHIM_ERR("KSPSetUp: operators not set"); return 0; }
static const char *ksp_reason_name(int r) { /* KSPConvergedReasons[] spellings */
  switch (r) {
  case 2: return "CONVERGED_RTOL";
  case 3: return "CONVERGED_ATOL";
  case 4: return "CONVERGED_ITS";
  case -3: return "DIVERGED_ITS";
  case -4: return "DIVERGED_DTOL";
  case -5: return "DIVERGED_BREAKDOWN";
  case -8: return "DIVERGED_INDEFINITE_PC";
  case -9: return "DIVERGED_NANORINF";
  default: return r > 0 ? "CONVERGED_UNKNOWN" : "DIVERGED_UNKNOWN";
  }
}
PetscErrorCode KSPSolve(KSP ksp, Vec b, Vec x) {
  if (!ksp->A || !ksp->A->bk) return SHIM_ERR("KSPSolve: operator not assembled");
  if (ksp->A != ksp->P) return SHIM_ERR("KSPSolve: Amat != Pmat is not supported by the shim (the reference passes A,A)");
  if (b->n != ksp->A->nrows || x->n != b->n) return SHIM_ERR("KSPSolve: size mismatch");
  char *opts = solver_options_text();
  int rc = shimbk_ksp_solve(ksp->A->bk, opts, b->n, b->a, x->a, &ksp->its, &ksp->reason, &ksp->rnorm);
  free(opts);
  if (rc) { fprintf(stderr, "[petsc-shim] back end '%s' error: %s\n", shimbk_name(), shimbk_last_error()); return 76; }
  if (opt_get("ksp_monitor")) { /* KSPMonitorResidual's line format; GMRES-type methods log the cycle start again at every restart */
    int len = 0;
    shimbk_ksp_history(NULL, 0, &len);
    double *h = (double *)malloc(sizeof(double) * (size_t)(len > 0 ? len : 1));
    shimbk_ksp_history(h, len, &len);
    const char *t = opt_get("ksp_type"), *rs = opt_get("ksp_gmres_restart");
    const int gmres_like = !t || !strcmp(t, "gmres") || !strcmp(t, "fgmres"); /* KSPGMRES is the default type */
    const int restart = rs ? atoi(rs) : 30;
    for (int i = 0; i < len; ++i) {
      const int it = gmres_like && restart > 0 ? (i / (restart + 1)) * restart + i % (restart + 1) : i;
      printf("%3d KSP Residual norm %14.12e \n", it, h[i]);
    }
    free(h);
  }
  if (opt_get("ksp_converged_reason"))
    printf("Linear solve %s due to %s iterations %d\n", ksp->reason > 0 ? "converged" : "did not converge", ksp_reason_name(ksp->reason), ksp->its);
  return 0;
}
PetscErrorCode KSPGetIterationNumber(KSP ksp, PetscInt *its) { *its = ksp->its; return 0; }
PetscErrorCode KSPGetConvergedReason(KSP ksp, PetscInt *reason) { *reason = ksp->reason; return 0; }
PetscErrorCode KSPGetResidualNorm(KSP ksp, PetscReal *rnorm) { *rnorm = ksp->rnorm; return 0; }
PetscErrorCode KSPDestroy(KSP *ksp) { if (ksp && *ksp) { free(*ksp); *ksp = NULL; } return 0; }
