/*
 * sp_oracle_amg.c -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * Aggregation multigrid set-up for an assembled block ("AMG V-cycle for A00 and for
 * L = B B^T in LSC"; PETSc analogue PCGAMG, -pc_gamg_type agg).  PARITY UNPINNED against PETSc: GAMG's aggregates
 * depend on a randomised greedy MIS and on its version, neither of which is available here, so this file DEFINES the
 * algorithm (deterministic, schedule-independent) and libb200sp's CUDA kernels must reproduce it: the aggregates and
 * the tentative prolongator bit for bit, everything after that to rounding.
 *
 *   node graph   : nodes i != j are neighbours when the bs x bs block (i,j) has strength s_ij = sum |a| > 0
 *                  (theta > 0:  s_ij > theta * sqrt(s_ii * s_jj)); nodes without neighbours (Dirichlet rows,
 *                  src/Discretization.c:268 leaves them as identity rows) are left out of the coarse space
 *   roots        : maximal independent set of distance 2 (Bell, Dalton, Olson: "Exposing fine-grained parallelism in
 *                  algebraic multigrid methods", SISC 2012, algorithm 5) with the priority key hash(i):i, or i itself
 *   aggregates   : root + its neighbours, then every remaining node joins the aggregate of its highest-key
 *                  aggregated neighbour; aggregates are numbered by ascending root index
 *   prolongator  : tentative P_t[(i,c),(agg(i),c)] = sqrt(w_i / sum of w over the aggregate), w_i = number of finest-level
 *                  nodes behind node i (the bs constant near-null-space vectors of the FINEST level, carried down
 *                  the hierarchy and orthonormalised per aggregate), smoothed once: P = P_t - omega D^-1 A P_t, omega = 4/(3 lambda)
 *   coarse matrix: Galerkin P^T A P
 */
#include "sp_oracle.h"
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static void *amalloc(size_t n) {
  void *p = malloc(n ? n : 1);
  if (!p) abort();
  return p;
}

/* priority of node i in the independent-set selection, distinct for distinct i and < 2^62.
 * order 0: hashed (a pseudo-random permutation: ~10 rounds whatever the numbering; aggregates of irregular size)
 * order 1: the node number itself (greedy in descending natural order: on a lexicographically numbered grid the roots
 *          form a regular lattice, 3 x 3 aggregates, at the price of O(grid side) rounds) */
static uint64_t amg_key(int i, int order) {
  if (order == 1) return (uint64_t)(uint32_t)i + 1;
  unsigned int h = (unsigned int)i * 2654435761u;
  h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13;
  return ((uint64_t)(h >> 2) << 32) | (uint32_t)i;
}

/* one merged pass over the bs rows of node i: calls back once per neighbouring node with the block strength */
typedef void (*amg_visit)(int j, double s, void *u);
static void amg_walk_node(const OrCsr *A, int bs, int i, amg_visit f, void *u) {
  int p[8], e[8];
  for (int c = 0; c < bs; ++c) { p[c] = A->rowptr[i * bs + c]; e[c] = A->rowptr[i * bs + c + 1]; }
  for (;;) {
    int j = -1;
    for (int c = 0; c < bs; ++c)
      if (p[c] < e[c]) { int n = A->col[p[c]] / bs; if (j < 0 || n < j) j = n; }
    if (j < 0) break;
    double s = 0.0;
    for (int c = 0; c < bs; ++c) /* row-major over the block */
      while (p[c] < e[c] && A->col[p[c]] / bs == j) { s += fabs(A->val[p[c]]); ++p[c]; }
    f(j, s, u);
  }
}
struct diag_ctx { int i; double s; };
static void visit_diag(int j, double s, void *u) { struct diag_ctx *d = (struct diag_ctx *)u; if (j == d->i) d->s = s; }
struct nbr_ctx { int i, n; int *out; const double *sd; double theta; };
static void visit_nbr(int j, double s, void *u) {
  struct nbr_ctx *d = (struct nbr_ctx *)u;
  if (j == d->i) return;
  int keep = d->theta > 0.0 ? (s > d->theta * sqrt(d->sd[d->i] * d->sd[j])) : (s > 0.0);
  if (keep) { if (d->out) d->out[d->n] = j; d->n++; }
}

/* strength graph on nodes; *grp (nn+1) and *gcol are malloc'ed */
static void amg_graph(const OrCsr *A, int bs, double theta, int **grp, int **gcol) {
  int nn = A->nrows / bs;
  double *sd = (double *)amalloc(sizeof(double) * (size_t)nn);
  for (int i = 0; i < nn; ++i) { struct diag_ctx d = {i, 0.0}; amg_walk_node(A, bs, i, visit_diag, &d); sd[i] = d.s; }
  int *rp = (int *)amalloc(sizeof(int) * ((size_t)nn + 1));
  rp[0] = 0;
  for (int i = 0; i < nn; ++i) { struct nbr_ctx d = {i, 0, NULL, sd, theta}; amg_walk_node(A, bs, i, visit_nbr, &d); rp[i + 1] = rp[i] + d.n; }
  int *gc = (int *)amalloc(sizeof(int) * (size_t)rp[nn]);
  for (int i = 0; i < nn; ++i) { struct nbr_ctx d = {i, 0, gc + rp[i], sd, theta}; amg_walk_node(A, bs, i, visit_nbr, &d); }
  free(sd);
  *grp = rp; *gcol = gc;
}

int or_amg_aggregate(const OrCsr *A, int bs, double theta, int order, int *agg) {
  int nn = A->nrows / bs;
  int *rp, *gc;
  amg_graph(A, bs, theta, &rp, &gc);
  /* t = state << 62 | key ; state 0 decided-out / left out, 1 undecided, 2 root */
  uint64_t *t = (uint64_t *)amalloc(8 * (size_t)nn), *m1 = (uint64_t *)amalloc(8 * (size_t)nn), *m2 = (uint64_t *)amalloc(8 * (size_t)nn);
  long undecided = 0;
  for (int i = 0; i < nn; ++i) {
    if (rp[i + 1] > rp[i]) { t[i] = ((uint64_t)1 << 62) | amg_key(i, order); ++undecided; }
    else t[i] = 0;
  }
  while (undecided) {
    for (int i = 0; i < nn; ++i) { uint64_t m = t[i]; for (int k = rp[i]; k < rp[i + 1]; ++k) if (t[gc[k]] > m) m = t[gc[k]]; m1[i] = m; }
    for (int i = 0; i < nn; ++i) { uint64_t m = m1[i]; for (int k = rp[i]; k < rp[i + 1]; ++k) if (m1[gc[k]] > m) m = m1[gc[k]]; m2[i] = m; }
    for (int i = 0; i < nn; ++i) {
      if ((t[i] >> 62) != 1) continue;
      if (m2[i] == t[i]) { t[i] = ((uint64_t)2 << 62) | amg_key(i, order); --undecided; }
      else if ((m2[i] >> 62) == 2) { t[i] = 0; --undecided; }
    }
  }
  /* number the roots, then the two joining passes (each reads only the previous pass) */
  int nagg = 0;
  int *a1 = (int *)amalloc(sizeof(int) * (size_t)nn);
  for (int i = 0; i < nn; ++i) a1[i] = (t[i] >> 62) == 2 ? nagg++ : -1;
  for (int i = 0; i < nn; ++i) agg[i] = a1[i];
  for (int i = 0; i < nn; ++i) {
    if (a1[i] >= 0 || rp[i + 1] == rp[i]) continue;
    uint64_t best = 0; int who = -1;
    for (int k = rp[i]; k < rp[i + 1]; ++k) { int j = gc[k]; if ((t[j] >> 62) == 2 && amg_key(j, order) >= best) { best = amg_key(j, order); who = j; } }
    if (who >= 0) agg[i] = a1[who];
  }
  memcpy(a1, agg, sizeof(int) * (size_t)nn);
  for (int i = 0; i < nn; ++i) {
    if (a1[i] >= 0 || rp[i + 1] == rp[i]) continue;
    uint64_t best = 0; int who = -1;
    for (int k = rp[i]; k < rp[i + 1]; ++k) { int j = gc[k]; if (a1[j] >= 0 && amg_key(j, order) >= best) { best = amg_key(j, order); who = j; } }
    if (who >= 0) agg[i] = a1[who];
  }
  free(a1); free(t); free(m1); free(m2); free(rp); free(gc);
  return nagg;
}

/* w[i] = number of finest-level nodes node i stands for (NULL: all 1): the constant vector on the finest level is
 * sqrt(w) on this one, so P_t[(i,c),(a,c)] = sqrt(w_i / W_a), W_a = sum of w over the aggregate (integers: the
 * result does not depend on the summation order); wc[a] = W_a (may be NULL) */
OrCsr *or_amg_tentative(int nn, int bs, const int *agg, int nagg, const int *w, int *wc) {
  int *W = (int *)calloc((size_t)nagg + 1, sizeof(int));
  long nnz = 0;
  for (int i = 0; i < nn; ++i) if (agg[i] >= 0) { W[agg[i]] += w ? w[i] : 1; nnz += bs; }
  OrCsr *P = or_csr_alloc(nn * bs, nagg * bs, nnz);
  long k = 0;
  for (int i = 0; i < nn; ++i)
    for (int c = 0; c < bs; ++c) {
      P->rowptr[i * bs + c] = (int)k;
      if (agg[i] >= 0) { P->col[k] = agg[i] * bs + c; P->val[k] = sqrt((double)(w ? w[i] : 1) / (double)W[agg[i]]); ++k; }
    }
  P->rowptr[nn * bs] = (int)k;
  if (wc) memcpy(wc, W, sizeof(int) * (size_t)nagg);
  free(W);
  return P;
}

OrCsr *or_csr_scale_rows(const OrCsr *A, const double *d) { /* diag(d) * A (copy) */
  long nnz = or_csr_nnz(A);
  OrCsr *C = or_csr_alloc(A->nrows, A->ncols, nnz);
  memcpy(C->rowptr, A->rowptr, sizeof(int) * ((size_t)A->nrows + 1));
  memcpy(C->col, A->col, sizeof(int) * (size_t)nnz);
  for (int i = 0; i < A->nrows; ++i)
    for (int k = A->rowptr[i]; k < A->rowptr[i + 1]; ++k) C->val[k] = d[i] * A->val[k];
  return C;
}

/* P = P_t - omega * D^-1 (A P_t), D = diag(A) with 0 -> 1 like PCJACOBI */
OrCsr *or_amg_smooth_prolongator(const OrCsr *A, const OrCsr *Pt, double omega) {
  double *d = (double *)amalloc(sizeof(double) * (size_t)A->nrows);
  or_csr_get_diagonal(A, d);
  for (int i = 0; i < A->nrows; ++i) d[i] = d[i] == 0.0 ? 1.0 : 1.0 / d[i];
  OrCsr *AP = or_csr_matmat(A, Pt);
  OrCsr *DAP = or_csr_scale_rows(AP, d);
  OrCsr *P = or_csr_add_scaled(Pt, -omega, DAP);
  or_csr_free(AP); or_csr_free(DAP); free(d);
  return P;
}

OrCsr *or_amg_galerkin(const OrCsr *A, const OrCsr *P) { /* P^T (A P) */
  OrCsr *R = or_csr_transpose(P);
  OrCsr *AP = or_csr_matmat(A, P);
  OrCsr *Ac = or_csr_matmat(R, AP);
  or_csr_free(R); or_csr_free(AP);
  return Ac;
}
