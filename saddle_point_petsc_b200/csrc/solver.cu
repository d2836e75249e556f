// solver.cu -- Krylov drivers (GMRES / FGMRES / MINRES / Chebyshev / Richardson), the fieldsplit-Schur,
// LSC and multigrid preconditioners, and the PETSc-style options wiring.  Host code only orchestrates:
// every vector or matrix operation is one of the hand-written kernels in kernels_*.cu; the host keeps
// the Hessenberg / Givens / Lanczos scalars exactly as PETSc does.
#include "solver.h"
#include "dist.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <sstream>

namespace b200sp {

static std::string pad(int n) { return std::string((size_t)n, ' '); }

// ------------------------------------------------------------------ context reductions
void Ctx::fetch_scalars(const double *d, int k, double *host) {
  B2_REQUIRE(k <= N_SCALARS, "fetch_scalars: too many scalars");
  B2_CUDA(cudaMemcpyAsync(h_scalars, d, sizeof(double) * (size_t)k, cudaMemcpyDeviceToHost, stream));
  B2_CUDA(cudaStreamSynchronize(stream));
  if (h_err && *h_err) { // sticky on the device side only until reported: clear it so that later calls are judged on their own
    const int code = *h_err;
    *h_err = 0;
    throw Error(B200SP_ERR_NCCL, "peer-to-peer halo exchange timed out waiting for a neighbour (code " + std::to_string(code) + ")");
  }
  std::memcpy(host, h_scalars, sizeof(double) * (size_t)k);
}
static void allreduce_sum(Ctx *c, double *d, int k) { // MPI_Allreduce(SUM) equivalent for VecDot/VecMDot/VecNorm
  if (!c->dcomm) return;
  if (c->profile) {
    cudaEventRecord(c->pev0, c->stream);
    c->reducer->allreduce_sum(d, k, c->stream);
    cudaEventRecord(c->pev1, c->stream);
    cudaEventSynchronize(c->pev1);
    float ms = 0;
    cudaEventElapsedTime(&ms, c->pev0, c->pev1);
    auto &pe = c->prof["comm:allreduce"];
    pe.ms += ms; pe.n++;
    return;
  }
  c->reducer->allreduce_sum(d, k, c->stream);
}

// ------------------------------------------------------------------ operators
std::string CsrOp::view(int indent) const {
  std::ostringstream o;
  o << pad(indent) << "Mat csr " << A->nrows << "x" << A->ncols << " nnz=" << A->nnz << " spmv_kernel=" << A->kernel << "\n";
  return o.str();
}

// Row-partitioned runs: A00 and A10 read the same x0, A01 and A11 the same x1, and all four blocks share the node
// halo, so each sub-vector is exchanged once (the second block of each pair reuses the ghosts).  The per-row
// summation order (A00 x0 first, then += A01 x1; A10 x0 first, then += A11 x1) is unchanged.
void NestOp::apply(const double *x, double *y) {
  const int64_t n0c = b00->ncols, n0r = b00->nrows;
  const bool share = b00->halo && b00->halo == b10->halo && b00->halo_dof == b10->halo_dof && (!b11 || (b01->halo == b11->halo && b01->halo_dof == b11->halo_dof));
  csr_spmv(*b00, x, y);
  csr_spmv(*b10, x, y + n0r, 1.0, nullptr, 0.0, share);
  csr_spmv(*b01, x + n0c, y, 1.0, y, 1.0);
  if (b11) csr_spmv(*b11, x + n0c, y + n0r, 1.0, y + n0r, 1.0, share);
}
void NestOp::residual(const double *b, const double *x, double *r) {
  const int64_t n0c = b00->ncols, n0r = b00->nrows;
  const bool share = b00->halo && b00->halo == b10->halo && b00->halo_dof == b10->halo_dof && (!b11 || (b01->halo == b11->halo && b01->halo_dof == b11->halo_dof));
  csr_spmv(*b00, x, r, -1.0, b, 1.0);
  csr_spmv(*b10, x, r + n0r, -1.0, b + n0r, 1.0, share);
  csr_spmv(*b01, x + n0c, r, -1.0, r, 1.0);
  if (b11) csr_spmv(*b11, x + n0c, r + n0r, -1.0, r + n0r, 1.0, share);
}
std::string NestOp::view(int indent) const {
  std::ostringstream o;
  o << pad(indent) << "Mat nest 2x2: A00 " << b00->nrows << "x" << b00->ncols << " nnz=" << b00->nnz << ", A01 nnz=" << b01->nnz
    << ", A10 nnz=" << b10->nnz << ", A11 nnz=" << (b11 ? b11->nnz : 0) << "\n";
  return o.str();
}

JacobiOp::JacobiOp(const Csr &A) : Op(A.ctx, A.nrows, A.nrows), dinv((size_t)A.nrows + 2) {
  B2_REQUIRE(A.nrows == A.ncols, "jacobi: matrix must be square");
  csr_get_diagonal(A, dinv.p);
  vec_reciprocal_safe(ctx, A.nrows, dinv.p); // zero diagonal -> 1 (PCJACOBI), stored as reciprocal (VecReciprocal)
}
std::string JacobiOp::view(int indent) const { return pad(indent) + "PC jacobi\n"; }

DenseInvOp::DenseInvOp(const Csr &A) : Op(A.ctx, A.nrows, A.nrows), Ainv((size_t)A.nrows * A.nrows) { dense_inverse_from_csr(A, Ainv.p); }
std::string DenseInvOp::view(int indent) const { return pad(indent) + "PC dense inverse (coarse LU stand-in) n=" + std::to_string(n_in) + "\n"; }

KspOp::KspOp(Ksp *k) : Op(k->ctx, k->n, k->n), ksp(k) {}
bool KspOp::capturable() const { // no convergence test (no host read-back) and a capturable preconditioner
  const bool fixed = ksp->type == KSP_PREONLY || ((ksp->type == KSP_CHEBYSHEV || ksp->type == KSP_RICHARDSON) && ksp->norm_none);
  return fixed && (!ksp->M || ksp->M->capturable()) && ksp->A->capturable();
}
bool MgOp::capturable() const {
  if (replicated && !(ctx->dcomm && ctx->dcomm->capturable() && replicated->capturable())) return false;
  for (auto &L : lev)
    if (L->smooth && !(L->smooth->norm_none && (L->smooth->type == KSP_CHEBYSHEV || L->smooth->type == KSP_RICHARDSON))) return false;
  return true;
}
void KspOp::apply(const double *x, double *y) { ksp->solve(x, y, false); }
std::string KspOp::view(int indent) const { return ksp->view(indent); }

SchurOp::SchurOp(std::shared_ptr<Csr> a11, std::shared_ptr<Csr> a10, Op *k0, std::shared_ptr<Csr> a01)
    : Op(a10->ctx, a10->nrows, a10->nrows), A11(a11), A10(a10), A01(a01), K0(k0), t0((size_t)a01->nrows + 2), t1((size_t)a01->nrows + 2) {}
void SchurOp::apply(const double *x, double *y) {
  csr_spmv(*A01, x, t0.p);
  K0->apply(t0.p, t1.p);
  if (A11) {
    csr_spmv(*A11, x, y);
    csr_spmv(*A10, t1.p, y, -1.0, y, 1.0); // y = A11 x - A10 t1
  } else {
    csr_spmv(*A10, t1.p, y, -1.0);
  }
}
std::string SchurOp::view(int indent) const {
  return pad(indent) + "Mat schurcomplement: S = A11 - A10 ksp(A00) A01, inner KSP:\n" + K0->view(indent + 2);
}

FieldSplitOp::FieldSplitOp(int fact_, double scale_, std::shared_ptr<Csr> a01, std::shared_ptr<Csr> a10, Op *k0, Op *ks)
    : Op(a01->ctx, a01->nrows + a10->nrows, a01->nrows + a10->nrows), fact(fact_), scale(scale_), A01(a01), A10(a10), K0(k0), KS(ks),
      t0((size_t)a01->nrows + 2), t1((size_t)a10->nrows + 2) {}
void FieldSplitOp::apply(const double *b, double *y) {
  const int64_t n0 = A01->nrows, n1 = A10->nrows;
  const double *b0 = b, *b1 = b + n0;
  double *y0 = y, *y1 = y + n0;
  switch (fact) {
  case 0: // DIAG: y0 = K0 b0 ; y1 = scale * KS b1
    K0->apply(b0, y0);
    KS->apply(b1, y1);
    vec_scale(ctx, n1, scale, y1);
    break;
  case 1: // LOWER
    K0->apply(b0, y0);
    csr_spmv(*A10, y0, t1.p, -1.0, b1, 1.0); // t1 = b1 - A10 y0
    KS->apply(t1.p, y1);
    break;
  case 2: // UPPER
    KS->apply(b1, y1);
    csr_spmv(*A01, y1, t0.p, -1.0, b0, 1.0); // t0 = b0 - A01 y1
    K0->apply(t0.p, y0);
    break;
  default: // FULL
    K0->apply(b0, y0);
    csr_spmv(*A10, y0, t1.p, -1.0, b1, 1.0);
    KS->apply(t1.p, y1);
    csr_spmv(*A01, y1, t0.p, -1.0, b0, 1.0);
    K0->apply(t0.p, y0);
    break;
  }
}
std::string FieldSplitOp::view(int indent) const {
  static const char *names[] = {"diag", "lower", "upper", "full"};
  std::ostringstream o;
  o << pad(indent) << "PC fieldsplit schur, factorization " << names[fact] << ", schur scale " << scale << "\n";
  o << pad(indent) << " split 0 (A00) solver:\n" << K0->view(indent + 2);
  o << pad(indent) << " split 1 (S) solver:\n" << KS->view(indent + 2);
  return o.str();
}

StridedSplitOp::StridedSplitOp(Op *in, const std::vector<int> &m) : Op(in->ctx, in->n_in, in->n_out), inner(in), map(m.size() + 1), xs(m.size() + 2), ys(m.size() + 2) {
  B2_CUDA(cudaMemcpyAsync(map.p, m.data(), sizeof(int) * m.size(), cudaMemcpyHostToDevice, ctx->stream));
  ctx->sync();
}
void StridedSplitOp::apply(const double *x, double *y) {
  vec_permute_gather(ctx, n_in, map.p, x, xs.p);   // VecScatter: monolithic -> [split 0; split 1]
  inner->apply(xs.p, ys.p);
  vec_permute_scatter(ctx, n_in, map.p, ys.p, y);  // and back
}
std::string StridedSplitOp::view(int indent) const { return pad(indent) + "strided fields (block size from the matrix), splits gathered/scattered\n" + inner->view(indent + 1); }

LscOp::LscOp(std::shared_ptr<Csr> a00, std::shared_ptr<Csr> a01, std::shared_ptr<Csr> a10, Op *linv, bool scale_diag_)
    : Op(a10->ctx, a10->nrows, a10->nrows), A00(a00), A01(a01), A10(a10), Linv(linv), scale_diag(scale_diag_),
      p0((size_t)a10->nrows + 2), p1((size_t)a10->nrows + 2), u0((size_t)a00->nrows + 2), u1((size_t)a00->nrows + 2) {
  if (scale_diag) {
    dinv.alloc((size_t)a00->nrows + 2);
    csr_get_diagonal(*A00, dinv.p);
    vec_reciprocal_safe(ctx, A00->nrows, dinv.p);
  }
}
void LscOp::apply(const double *x, double *y) { // y = Linv A10 [D^-1] A00 [D^-1] A01 Linv x
  const int64_t n0 = A00->nrows;
  Linv->apply(x, p0.p);
  csr_spmv(*A01, p0.p, u0.p);
  if (scale_diag) vec_pointwise_mult(ctx, n0, u0.p, dinv.p, u0.p);
  csr_spmv(*A00, u0.p, u1.p);
  if (scale_diag) vec_pointwise_mult(ctx, n0, u1.p, dinv.p, u1.p);
  csr_spmv(*A10, u1.p, p1.p);
  Linv->apply(p1.p, y);
}
std::string LscOp::view(int indent) const {
  return pad(indent) + "PC lsc" + (scale_diag ? " (scale_diag)" : "") + ", L = A10 A01 solver:\n" + Linv->view(indent + 2);
}

// PCMG multiplicative V-cycle: smoothdown (zero guess), residual, restrict, recurse, interpolate-add, smoothup
void MgOp::cycle(int l, const double *b, double *x) {
  Level &L = *lev[l];
  const bool last = l == (int)lev.size() - 1;
  if (last && !replicated) { coarse->apply(b, x); return; }
  // Row-partitioned levels: every vector below is consumed by a known matrix, so its producer pushes the halo
  // (smoother -> A for the residual, residual -> R, interpolation -> A for the post-smoother, post-smoother -> the
  // finer level's P); without peer-to-peer halos the hints are ignored.
  Halo *hA = L.A->halo.get(), *hR = L.R ? L.R->halo.get() : nullptr;
  L.smooth->push_after = hA; L.smooth->push_after_dof = L.A->halo_dof;
  L.smooth->solve(b, x, false);
  csr_spmv(*L.A, x, L.r.p, -1.0, b, 1.0, false, hR, hR ? L.R->halo_dof : 0); // r = b - A x
  if (!last) {
    Level &Lc = *lev[l + 1];
    csr_spmv(*L.R, L.r.p, Lc.b.p);                  // restrict
    cycle(l + 1, Lc.b.p, Lc.x.p);
    csr_spmv(*L.P, Lc.x.p, x, 1.0, x, 1.0, false, hA, L.A->halo_dof); // x += P xc
  } else {
    // bridge to the replicated coarse hierarchy: restrict into my part of the coarse vector, all-gather, reorder to
    // the natural numbering, run the remaining levels redundantly on every rank, take my part back, interpolate
    csr_spmv(*L.R, L.r.p, loc_b.p);
    if (ctx->profile) cudaEventRecord(ctx->pev0, ctx->stream);
    bridge->allgather(loc_b.p, g_all.p, bridge_cnt, ctx->stream);
    if (ctx->profile) {
      cudaEventRecord(ctx->pev1, ctx->stream);
      cudaEventSynchronize(ctx->pev1);
      float ms = 0;
      cudaEventElapsedTime(&ms, ctx->pev0, ctx->pev1);
      auto &pe = ctx->prof["comm:allgather"];
      pe.ms += ms; pe.n++;
    }
    vec_permute_scatter(ctx, (int64_t)bridge_cnt * ctx->size, gather_map.p, g_all.p, nat_b.p);
    replicated->apply(nat_b.p, nat_x.p);
    vec_permute_gather(ctx, bridge_nloc, local_map.p, nat_x.p, loc_x.p);
    csr_spmv(*L.P, loc_x.p, x, 1.0, x, 1.0, false, hA, L.A->halo_dof);
  }
  Halo *hUp = l > 0 && lev[l - 1]->P ? lev[l - 1]->P->halo.get() : nullptr; // the finer level interpolates this level's x
  L.smooth->push_after = hUp; L.smooth->push_after_dof = hUp ? lev[l - 1]->P->halo_dof : 0;
  L.smooth->solve(b, x, true);
  L.smooth->push_after = nullptr;
}
void MgOp::apply(const double *b, double *x) { cycle(0, b, x); } // level 0 works on the caller's vectors: no copies
std::string MgOp::view(int indent) const {
  std::ostringstream o;
  o << pad(indent) << "PC mg: multiplicative V-cycle, " << lev.size() << " levels (" << kind << ")\n";
  for (size_t l = 0; l < lev.size(); ++l) {
    o << pad(indent + 1) << "level " << l << ": n=" << lev[l]->A->nrows << " nnz=" << lev[l]->A->nnz << "\n";
    if (lev[l]->smooth) o << lev[l]->smooth->view(indent + 3);
  }
  o << coarse->view(indent + 1);
  return o.str();
}

// ------------------------------------------------------------------ KSP
Ksp::~Ksp() {
  for (auto &kv : pc_graphs)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
}

// PCApply.  For the outer KSP the whole preconditioner (fieldsplit + multigrid V-cycle: several hundred short kernels
// and, on >1 GPU, the NCCL halo exchanges between them) is recorded once per (x, y) pair into a CUDA graph and
// replayed: the launch-latency-bound coarse levels then cost one graph launch instead of one launch each.
void Ksp::pc_apply(const double *x, double *y) {
  if (!M) { vec_copy(ctx, n, x, y); return; }
  if (!use_pc_graph || ctx->profile) { M->apply(x, y); return; }
  if (!pc_warmed) { M->apply(x, y); pc_warmed = true; return; } // first call runs eagerly: lazy allocations, func attributes
  auto key = std::make_pair(x, y);
  auto it = pc_graphs.find(key);
  if (it == pc_graphs.end()) {
    if (pc_graphs.size() >= 96) { M->apply(x, y); return; }
    cudaGraph_t graph = nullptr;
    const int64_t l0 = ctx->launches;
    B2_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    try {
      M->apply(x, y);
    } catch (...) {
      cudaStreamEndCapture(ctx->stream, &graph);
      if (graph) cudaGraphDestroy(graph);
      use_pc_graph = false;
      throw;
    }
    B2_CUDA(cudaStreamEndCapture(ctx->stream, &graph));
    PcGraph g;
    g.launches = ctx->launches - l0;
    B2_CUDA(cudaGraphInstantiate(&g.exec, graph, 0));
    B2_CUDA(cudaGraphDestroy(graph));
    it = pc_graphs.emplace(key, g).first;
  } else {
    ctx->launches += it->second.launches; // the recorded kernels run again
  }
  B2_CUDA(cudaGraphLaunch(it->second.exec, ctx->stream));
}
double Ksp::dot(const double *x, const double *y) {
  double r;
  vec_dot(ctx, n, x, y, ctx->d_scalars);
  allreduce_sum(ctx, ctx->d_scalars, 1);
  ctx->fetch_scalars(ctx->d_scalars, 1, &r);
  return r;
}
double Ksp::norm2(const double *x) { return std::sqrt(dot(x, x)); }

// KSPConvergedDefault
int Ksp::converged(int it, double rn) {
  if (keep_history) hist.push_back(rn);
  rnorm = rn;
  if (it == 0) rnorm0 = rn;
  if (std::isnan(rn) || std::isinf(rn)) return B200SP_DIVERGED_NANORINF;
  const double ttol = std::max(rtol * rnorm0, atol);
  if (rn <= ttol) return rn < atol ? B200SP_CONVERGED_ATOL : B200SP_CONVERGED_RTOL;
  if (rn >= dtol * rnorm0) return B200SP_DIVERGED_DTOL;
  return 0;
}

int Ksp::solve(const double *b, double *x, bool guess_nonzero) {
  B2_REQUIRE(A, "KSP: operators not set");
  its = 0;
  hist.clear();
  ld = (n + 15) & ~(int64_t)15;
  if (!guess_nonzero && type != KSP_PREONLY && type != KSP_CHEBYSHEV) vec_set(ctx, n, 0.0, x); // VecSet(x,0): zero initial guess
  switch (type) {
  case KSP_PREONLY:
    pc_apply(b, x);
    its = 1;
    reason = B200SP_CONVERGED_ITS;
    return reason;
  case KSP_RICHARDSON: return solve_richardson(b, x, guess_nonzero);
  case KSP_CHEBYSHEV: return solve_chebyshev(b, x, guess_nonzero);
  case KSP_GMRES: return solve_gmres(b, x, guess_nonzero, false);
  case KSP_FGMRES: return solve_gmres(b, x, guess_nonzero, true);
  case KSP_MINRES: return solve_minres(b, x, guess_nonzero);
  }
  throw Error(B200SP_ERR_UNSUPPORTED, "KSP: unknown type");
}

int Ksp::solve_richardson(const double *b, double *x, bool guess_nonzero) {
  if (!w0.p) { w0.alloc((size_t)ld); w1.alloc((size_t)ld); }
  double *r = w0.p, *z = w1.p;
  reason = 0;
  for (int it = 0; it < max_it; ++it) {
    const double *res = r;
    if (it == 0 && !guess_nonzero) res = b; else A->residual(b, x, r);
    const double *dinv = M ? M->jacobi_dinv() : nullptr;
    if (norm_none && (dinv || !M)) { // fused: x += scale * (dinv .* r)
      vec_cheb_update(ctx, n, 1.0, x, 0.0, x, richardson_scale, dinv, res, x);
    } else {
      pc_apply(res, z);
      if (!norm_none) {
        reason = converged(it, norm2(z));
        if (reason) { its = it; break; }
      }
      vec_axpy(ctx, n, richardson_scale, z, x);
    }
    its = it + 1;
  }
  if (!reason) reason = norm_none ? B200SP_CONVERGED_ITS : B200SP_DIVERGED_ITS;
  return reason;
}

// KSPSolve_Chebyshev recurrence; with a Jacobi PC the PC application is fused into the update
int Ksp::solve_chebyshev(const double *b, double *x, bool guess_nonzero) {
  if (!w0.p) { w0.alloc((size_t)ld); w1.alloc((size_t)ld); w2.alloc((size_t)ld); w3.alloc((size_t)ld); }
  B2_REQUIRE(emax > 0.0, "chebyshev: eigenvalue bounds not set");
  double *r = w0.p, *pkm1 = w1.p, *pk = w2.p, *pkp1 = w3.p;
  const double scale = 2.0 / (emax + emin), alpha = 1.0 - scale * emin, mu = 1.0 / alpha, omegaprod = 2.0 / alpha;
  double ckm1 = 1.0, ck = mu, ckp1, omega;
  const double *dinv = M ? M->jacobi_dinv() : nullptr;
  const bool fused = (dinv || !M) && norm_none;
  reason = 0;
  const double *res = b;
  if (guess_nonzero && !fused) { A->residual(b, x, r); res = r; }
  if (fused) {
    // Smoother mode (fixed sweeps, Jacobi or no PC): no copies and no zero-fill.  The previous iterate of the
    // first update is x itself (nonzero guess) or the zero vector (coefficient 0 on a finite dummy operand);
    // iterates ping-pong between two scratch vectors and the LAST update writes straight into x.  Every
    // update is elementwise, so writing over the vector that holds p_{k-1} is safe.
    // When the operator is a CSR matrix each sweep after the first is ONE kernel: the SpMV's epilogue forms the
    // residual, applies Jacobi and the three-term update (SpmvEpi::cheb) -- r is never written to memory.
    const Csr *Ac = A->csr();
    // `more`: another sweep multiplies `out` by the same matrix next; otherwise the caller's hint (push_after) applies
    auto sweep = [&](const double *pm1_, double ca, const double *cur_, double cb, double cc, double *out, bool more) {
      if (Ac) {
        SpmvEpi e;
        e.cheb = 1; e.z = b; e.pm1 = pm1_; e.pk = cur_; e.dinv = dinv; e.ca = ca; e.cb = cb; e.cc = cc;
        Halo *pt = more ? Ac->halo.get() : push_after;
        csr_spmv_epi(*Ac, cur_, out, e, false, pt, more ? Ac->halo_dof : push_after_dof);
      } else {
        A->residual(b, cur_, r);
        vec_cheb_update(ctx, n, ca, pm1_, cb, cur_, cc, dinv, r, out);
      }
    };
    const double *pm1 = guess_nonzero ? x : b; // b is only a finite dummy when the guess is zero
    double am1 = guess_nonzero ? 1.0 : 0.0;
    double *cur;
    if (guess_nonzero) {
      cur = w2.p;                                   // p1 = x0 + scale * M^-1 (b - A x0); never in place: the SpMV gathers x0
      if (max_it == 1) { // the copy below changes the vector the consumer sees: no push hint
        Halo *keep = push_after; push_after = nullptr;
        sweep(x, 1.0, x, 0.0, scale, cur, false);
        push_after = keep;
        vec_copy(ctx, n, cur, x);
      } else {
        sweep(x, 1.0, x, 0.0, scale, cur, true);
      }
    } else {
      cur = max_it == 1 ? x : w2.p;
      vec_cheb_update(ctx, n, 0.0, b, 0.0, b, scale, dinv, b, cur); // p1 = scale * M^-1 b
    }
    its = 1;
    for (int i = 1; i < max_it; ++i) {
      ckp1 = 2.0 * mu * ck - ckm1;
      omega = omegaprod * ck / ckp1;
      double *out = (i == max_it - 1) ? x : (cur == w2.p ? w3.p : w2.p);
      sweep(pm1, (1.0 - omega) * am1, cur, omega, omega * scale, out, i < max_it - 1);
      pm1 = cur; am1 = 1.0; cur = out;
      ckm1 = ck; ck = ckp1;
      its = i + 1;
    }
    reason = B200SP_CONVERGED_ITS;
    return reason;
  }
  if (!guess_nonzero) vec_set(ctx, n, 0.0, x);
  vec_copy(ctx, n, x, pkm1);
  if (fused) {
    vec_cheb_update(ctx, n, 1.0, pkm1, 0.0, pkm1, scale, dinv, res, pk); // pk = x + scale * M^-1 r
  } else {
    pc_apply(res, pk);
    if (!norm_none) reason = converged(0, norm2(pk));
    vec_aypx(ctx, n, scale, pkm1, pk);
  }
  its = 1;
  for (int i = 1; i < max_it && !reason; ++i) {
    A->residual(b, pk, r);
    ckp1 = 2.0 * mu * ck - ckm1;
    omega = omegaprod * ck / ckp1;
    if (fused) {
      vec_cheb_update(ctx, n, 1.0 - omega, pkm1, omega, pk, omega * scale, dinv, r, pkp1);
    } else {
      pc_apply(r, pkp1);
      if (!norm_none) { reason = converged(i, norm2(pkp1)); if (reason) break; }
      vec_axpbypcz(ctx, n, 1.0 - omega, pkm1, omega, pk, omega * scale, pkp1, pkp1);
    }
    double *t = pkm1; pkm1 = pk; pk = pkp1; pkp1 = t;
    ckm1 = ck; ck = ckp1;
    its = i + 1;
  }
  vec_copy(ctx, n, pk, x);
  if (!reason) reason = norm_none ? B200SP_CONVERGED_ITS : B200SP_DIVERGED_ITS;
  return reason;
}

// KSPSolve_GMRES / KSPSolve_FGMRES: classical Gram-Schmidt without refinement.  Per iteration:
//   SpMV(+PC)  ->  k_mdot (h = V^T w)  ->  k_maxpy+norm (w -= V h, ||w||^2)  ->  k_scale_inv_sqrt  ->  ONE host sync
// h never leaves the device between the three vector kernels.
int Ksp::solve_gmres(const double *b, double *x, bool guess_nonzero, bool flexible) {
  const int m = restart;
  B2_REQUIRE(m >= 1 && 2 * m + 4 <= N_SCALARS, "gmres: restart must be in [1,62]");
  if (!V.p) {
    V.alloc((size_t)ld * (m + 1));
    if (flexible) Z.alloc((size_t)ld * m);
    w0.alloc((size_t)ld);
  }
  std::vector<double> H((size_t)(m + 1) * m, 0.0), cs(m), sn(m), g(m + 1), y(m), hcol(m + 2);
  double *d_h = ctx->d_scalars;
  int itc = 0;
  bool first = true;
  reason = 0;
  while (!reason) {
    double *v0 = V.p;
    if (first && !guess_nonzero) {
      if (flexible || !M) vec_copy(ctx, n, b, v0); else pc_apply(b, v0);
    } else {
      if (flexible || !M) A->residual(b, x, v0);
      else { A->residual(b, x, w0.p); pc_apply(w0.p, v0); }
    }
    first = false;
    const double beta = norm2(v0);
    reason = converged(itc, beta);
    if (reason) break;
    if (itc >= max_it) { reason = B200SP_DIVERGED_ITS; break; }
    vec_scale(ctx, n, 1.0 / beta, v0);
    g[0] = beta;
    int it = 0;
    while (!reason && it < m && itc < max_it) {
      double *vk = V.p + (size_t)ld * it, *vn = V.p + (size_t)ld * (it + 1);
      if (flexible) {
        double *zk = Z.p + (size_t)ld * it;
        pc_apply(vk, zk);
        A->apply(zk, vn);
      } else if (M) {
        A->apply(vk, w0.p);
        pc_apply(w0.p, vn);
      } else {
        A->apply(vk, vn);
      }
      double *h = H.data() + (size_t)(m + 1) * it;
      if (orthog == ORTHOG_MGS) {
        // KSPGMRESModifiedGramSchmidtOrthogonalization: one vector at a time (h_j on the device between the dot
        // and the update; the norm of the last update is the Hessenberg sub-diagonal)
        for (int j = 0; j <= it; ++j) {
          vec_mdot(ctx, n, 1, vn, V.p + (size_t)ld * j, ld, d_h + j);
          allreduce_sum(ctx, d_h + j, 1);
          vec_maxpy_norm2(ctx, n, 1, vn, V.p + (size_t)ld * j, ld, d_h + j, d_h + it + 1);
        }
        allreduce_sum(ctx, d_h + it + 1, 1);
        ctx->fetch_scalars(d_h, it + 2, hcol.data());
      } else {
        vec_mdot(ctx, n, it + 1, vn, V.p, ld, d_h);                          // VecMDot
        allreduce_sum(ctx, d_h, it + 1);
        vec_maxpy_norm2(ctx, n, it + 1, vn, V.p, ld, d_h, d_h + it + 1);     // VecMAXPY + VecNorm fused
        allreduce_sum(ctx, d_h + it + 1, 1);
        ctx->fetch_scalars(d_h, it + 2, hcol.data());
        bool refine = orthog == ORTHOG_CGS_REFINE_ALWAYS;
        if (orthog == ORTHOG_CGS_REFINE_IFNEEDED) { // refine when what is left of w is smaller than what was removed
          double hnrm = 0.0;
          for (int j = 0; j <= it; ++j) hnrm += hcol[j] * hcol[j];
          refine = hcol[it + 1] < hnrm;
        }
        if (refine) { // second classical Gram-Schmidt pass; the corrections add to the Hessenberg column
          double *d_h2 = d_h + m + 2;
          vec_mdot(ctx, n, it + 1, vn, V.p, ld, d_h2);
          allreduce_sum(ctx, d_h2, it + 1);
          vec_maxpy_norm2(ctx, n, it + 1, vn, V.p, ld, d_h2, d_h + it + 1);
          allreduce_sum(ctx, d_h + it + 1, 1);
          std::vector<double> h2((size_t)it + 1);
          ctx->fetch_scalars(d_h2, it + 1, h2.data());
          ctx->fetch_scalars(d_h + it + 1, 1, &hcol[(size_t)it + 1]);
          for (int j = 0; j <= it; ++j) hcol[j] += h2[(size_t)j];
        }
      }
      vec_scale_inv_sqrt(ctx, n, d_h + it + 1, vn, vn);                    // v_{k+1} = w / ||w||
      for (int j = 0; j <= it; ++j) h[j] = hcol[j];
      const double hn = std::sqrt(hcol[it + 1]);
      h[it + 1] = hn;
      for (int j = 0; j < it; ++j) {
        const double a = h[j], bb = h[j + 1];
        h[j] = cs[j] * a + sn[j] * bb;
        h[j + 1] = -sn[j] * a + cs[j] * bb;
      }
      const double tt = std::sqrt(h[it] * h[it] + h[it + 1] * h[it + 1]);
      if (tt == 0.0 || std::isnan(tt)) { reason = std::isnan(tt) ? B200SP_DIVERGED_NANORINF : B200SP_DIVERGED_BREAKDOWN; break; }
      cs[it] = h[it] / tt;
      sn[it] = h[it + 1] / tt;
      g[it + 1] = -sn[it] * g[it];
      g[it] = cs[it] * g[it];
      h[it] = cs[it] * h[it] + sn[it] * h[it + 1];
      h[it + 1] = 0.0;
      const double res = std::fabs(g[it + 1]);
      it++;
      itc++;
      reason = converged(itc, res);
      if (!reason && hn == 0.0) reason = B200SP_DIVERGED_BREAKDOWN;
    }
    // KSPGMRESBuildSoln: R y = g, x += V y (Z y for FGMRES)
    if (it > 0) {
      for (int i = it - 1; i >= 0; --i) {
        double s = g[i];
        for (int j = i + 1; j < it; ++j) s -= H[i + (size_t)(m + 1) * j] * y[j];
        y[i] = s / H[i + (size_t)(m + 1) * i];
      }
      std::memcpy(ctx->h_scalars, y.data(), sizeof(double) * (size_t)it);
      B2_CUDA(cudaMemcpyAsync(d_h, ctx->h_scalars, sizeof(double) * (size_t)it, cudaMemcpyHostToDevice, ctx->stream));
      if (!flexible && M == nullptr) vec_maxpy(ctx, n, it, x, V.p, ld, d_h);
      else vec_maxpy(ctx, n, it, x, flexible ? Z.p : V.p, ld, d_h);
      ctx->sync(); // h_scalars is reused by the next fetch
    }
    if (!reason && itc >= max_it) reason = B200SP_DIVERGED_ITS;
  }
  its = itc;
  return reason;
}

// KSPSolve_MINRES (classic Paige-Saunders form, PETSc <= 3.18)
int Ksp::solve_minres(const double *b, double *x, bool guess_nonzero) {
  if (!w0.p) {
    DevBuf<double> *bufs[] = {&w0, &w1, &w2, &w3, &w4, &w5, &w6, &w7, &w8};
    for (auto *p : bufs) { p->alloc((size_t)ld); p->zero(ctx->stream); }
  }
  double *r = w0.p, *v = w1.p, *vold = w2.p, *z = w3.p, *u = w4.p, *uold = w5.p, *w = w6.p, *wold = w7.p, *wooold = w8.p;
  vec_set(ctx, n, 0.0, vold); vec_set(ctx, n, 0.0, uold); vec_set(ctx, n, 0.0, w); vec_set(ctx, n, 0.0, wold);
  double alpha, beta, betaold, eta, c = 1.0, cold = 1.0, s = 0.0, sold = 0.0, coold, soold, rho0, rho1, rho2, rho3, dp;
  int itc = 0;
  reason = 0;
  if (guess_nonzero) A->residual(b, x, r); else vec_copy(ctx, n, b, r);
  pc_apply(r, z);
  dp = dot(r, z);
  const double haptol = 1e-18; // KSPMINRES haptol: a tiny negative r.z is rounding, not an indefinite PC
  if (dp < 0.0 && std::fabs(dp) > haptol) { reason = B200SP_DIVERGED_INDEFINITE_PC; its = 0; return reason; }
  beta = std::sqrt(std::fabs(dp));
  eta = beta;
  dp = norm2(z);
  reason = converged(0, dp);
  if (reason) { its = 0; return reason; }
  if (beta == 0.0) { reason = B200SP_CONVERGED_ATOL; its = 0; return reason; }
  // v = r/beta, u = z/beta (in place: r,z buffers become v,u; the old v,u buffers are recycled as r,z)
  vec_scale(ctx, n, 1.0 / beta, r); vec_scale(ctx, n, 1.0 / beta, z);
  std::swap(r, v); std::swap(z, u);
  while (itc < max_it) {
    A->apply(u, r);
    alpha = dot(u, r);
    pc_apply(r, z);
    vec_axpbypcz(ctx, n, -alpha, v, -beta, vold, 1.0, r, r); // r -= alpha v + beta v_old
    vec_axpbypcz(ctx, n, -alpha, u, -beta, uold, 1.0, z, z); // z -= alpha u + beta u_old
    betaold = beta;
    { const double d = dot(r, z); if (d < 0.0 && std::fabs(d) > haptol) { reason = B200SP_DIVERGED_INDEFINITE_PC; break; } beta = std::sqrt(std::fabs(d)); }
    coold = cold; cold = c; soold = sold; sold = s;
    rho0 = cold * alpha - coold * sold * betaold;
    rho1 = std::sqrt(rho0 * rho0 + beta * beta);
    rho2 = sold * alpha + coold * cold * betaold;
    rho3 = soold * betaold;
    c = rho0 / rho1; s = beta / rho1;
    { // w_new = (u - rho2 w - rho3 w_old)/rho1, written over the w_oold buffer, then rotate
      const double irho1 = 1.0 / rho1;
      vec_axpbypcz(ctx, n, irho1, u, -rho2 * irho1, w, -rho3 * irho1, wold, wooold);
      double *t = wooold; wooold = wold; wold = w; w = t;
    }
    vec_axpy(ctx, n, c * eta, w, x);
    eta = -s * eta;
    // v_old <- v, v <- r/beta ; u_old <- u, u <- z/beta (buffer rotation, one scale each)
    if (beta != 0.0) { vec_scale(ctx, n, 1.0 / beta, r); vec_scale(ctx, n, 1.0 / beta, z); }
    { double *t = vold; vold = v; v = r; r = t; }
    { double *t = uold; uold = u; u = z; z = t; }
    dp = std::fabs(s) * dp;
    itc++;
    reason = converged(itc, dp);
    if (reason) break;
  }
  if (!reason) reason = B200SP_DIVERGED_ITS;
  its = itc;
  return reason;
}

std::string Ksp::view(int indent) const {
  static const char *names[] = {"preonly", "richardson", "chebyshev", "gmres", "fgmres", "minres"};
  std::ostringstream o;
  o << pad(indent) << "KSP (" << (prefix.empty() ? "outer" : prefix) << ") type " << names[type];
  if (type == KSP_GMRES || type == KSP_FGMRES) o << " restart=" << restart;
  if (type == KSP_CHEBYSHEV) o << " eigs=(" << emin << "," << emax << ")";
  if (type != KSP_PREONLY) o << " max_it=" << max_it << (norm_none ? " norm=none" : "") << " rtol=" << rtol;
  o << "\n";
  if (M) o << M->view(indent + 1); else o << pad(indent + 1) << "PC none\n";
  return o.str();
}

double estimate_lambda_max(Ctx *c, Op *A, Op *M, int nits, bool local_only) {
  const int64_t n = A->n_in;
  DevBuf<double> v((size_t)n + 2), t((size_t)n + 2), z((size_t)n + 2);
  // start vector: hashed by the NATURAL global index of every entry, so that the estimate (and with it the Chebyshev
  // coefficients, i.e. the preconditioner) is the same on any number of ranks
  const Csr *ac = A->csr();
  if (!local_only && ac && ac->halo && ac->layout && ac->dof_r > 0 && ac->nrows == ac->ncols) {
    int xs, ys, xm, ym;
    ac->layout->box(c->rank, &xs, &ys, &xm, &ym);
    B2_REQUIRE((int64_t)xm * ym * ac->dof_r == n, "estimate_lambda_max: matrix does not match its DMDA layout");
    vec_hash_natural(c, xs, ys, xm, ym, ac->layout->M, ac->dof_r, v.p);
  } else {
    vec_hash(c, n, v.p);
  }
  double lam = 0.0, nv;
  for (int it = 0; it < nits; ++it) {
    vec_dot(c, n, v.p, v.p, c->d_scalars);
    if (!local_only) allreduce_sum(c, c->d_scalars, 1);
    c->fetch_scalars(c->d_scalars, 1, &nv);
    vec_scale(c, n, 1.0 / std::sqrt(nv), v.p);
    A->apply(v.p, t.p);
    if (M) M->apply(t.p, z.p); else vec_copy(c, n, t.p, z.p);
    vec_dot(c, n, z.p, z.p, c->d_scalars);
    if (!local_only) allreduce_sum(c, c->d_scalars, 1);
    c->fetch_scalars(c->d_scalars, 1, &lam);
    lam = std::sqrt(lam);
    vec_copy(c, n, z.p, v.p);
  }
  c->sync();
  return lam;
}

// ------------------------------------------------------------------ options wiring
static bool is_number(const std::string &t) {
  char *end = nullptr;
  std::strtod(t.c_str(), &end);
  return end && *end == 0 && !t.empty();
}
void Solver::set_options(const char *text) {
  std::istringstream in(text ? text : "");
  std::vector<std::string> tok;
  std::string t;
  while (in >> t) tok.push_back(t);
  for (size_t i = 0; i < tok.size();) {
    B2_REQUIRE(tok[i].size() > 1 && tok[i][0] == '-', "options: expected -name, got '" + tok[i] + "'");
    std::string key = tok[i].substr(1);
    if (i + 1 < tok.size() && !(tok[i + 1][0] == '-' && !is_number(tok[i + 1]))) { opts[key] = tok[i + 1]; i += 2; }
    else { opts[key] = ""; i += 1; }
  }
  is_setup = false;
}
std::string Solver::opt(const std::string &key, const std::string &def) const {
  used[key] = true;
  auto it = opts.find(key);
  return it == opts.end() ? def : it->second;
}

// PETSc prints unused options with -options_left; here an unused SOLVER option is an error: it means a mistyped key or a
// feature this library does not have, and either way the solve would silently differ from what PETSc would run.
// Options the caller handles itself (monitors, viewers) are exempt.
void Solver::check_options_left() const {
  static const char *caller_side[] = {"ksp_monitor", "ksp_monitor_true_residual", "ksp_converged_reason", "ksp_view", "ksp_view_pre", "pc_view", nullptr};
  std::string left;
  for (auto &kv : opts) {
    const std::string &k = kv.first;
    const bool solver_key = k.find("ksp_") != std::string::npos || k.find("pc_") != std::string::npos || k.rfind("fieldsplit_", 0) == 0 ||
                            k.rfind("mg_", 0) == 0 || k.rfind("b200sp_", 0) == 0;
    if (!solver_key || used.count(k)) continue;
    bool exempt = false;
    for (const char **c = caller_side; *c; ++c) exempt = exempt || k == *c;
    if (!exempt) left += " -" + k;
  }
  if (!left.empty()) throw Error(B200SP_ERR_UNSUPPORTED, "options not used by any solver object (unknown or unsupported here):" + left);
}

static int ksp_type_from(const std::string &s) {
  if (s == "preonly") return KSP_PREONLY;
  if (s == "richardson") return KSP_RICHARDSON;
  if (s == "chebyshev") return KSP_CHEBYSHEV;
  if (s == "gmres") return KSP_GMRES;
  if (s == "fgmres") return KSP_FGMRES;
  if (s == "minres") return KSP_MINRES;
  throw Error(B200SP_ERR_UNSUPPORTED, "unsupported -ksp_type " + s);
}

Ksp *Solver::make_ksp(const std::string &prefix, Op *A, Op *M, const char *default_type) {
  ksps.emplace_back(new Ksp(ctx, prefix));
  Ksp *k = ksps.back().get();
  k->set_operators(A, M);
  const std::string tname = opt(prefix + "ksp_type", default_type);
  k->type = ksp_type_from(tname);
  k->rtol = std::stod(opt(prefix + "ksp_rtol", "1e-5"));
  k->atol = std::stod(opt(prefix + "ksp_atol", "1e-50"));
  k->dtol = std::stod(opt(prefix + "ksp_divtol", "1e5"));
  k->max_it = std::stoi(opt(prefix + "ksp_max_it", "10000"));
  k->restart = std::stoi(opt(prefix + "ksp_gmres_restart", "30"));
  k->richardson_scale = std::stod(opt(prefix + "ksp_richardson_scale", "1.0"));
  if (has(prefix + "ksp_gmres_modifiedgramschmidt")) k->orthog = ORTHOG_MGS;
  else {
    const std::string rt = opt(prefix + "ksp_gmres_cgs_refinement_type", "refine_never");
    B2_REQUIRE(rt == "refine_never" || rt == "refine_ifneeded" || rt == "refine_always", "bad -ksp_gmres_cgs_refinement_type " + rt);
    k->orthog = rt == "refine_always" ? ORTHOG_CGS_REFINE_ALWAYS : rt == "refine_ifneeded" ? ORTHOG_CGS_REFINE_IFNEEDED : ORTHOG_CGS;
  }
  const std::string nt = opt(prefix + "ksp_norm_type", "");
  // inner chebyshev / richardson solvers are fixed-sweep smoothers unless a norm type is requested
  if (nt == "none" || (nt.empty() && !prefix.empty() && (k->type == KSP_CHEBYSHEV || k->type == KSP_RICHARDSON))) k->norm_none = true;
  { // -ksp_pc_side / -ksp_norm_type: each method is implemented with PETSc's DEFAULT side and norm only; anything else is rejected
    const bool right = k->type == KSP_FGMRES;
    const std::string side = opt(prefix + "ksp_pc_side", right ? "right" : "left");
    if (side != (right ? "right" : "left"))
      throw Error(B200SP_ERR_UNSUPPORTED, "-" + prefix + "ksp_pc_side " + side + ": " + tname + " is implemented with " + (right ? "right" : "left") + " preconditioning only");
    const std::string natural = right ? "unpreconditioned" : "preconditioned";
    const bool krylov = k->type == KSP_GMRES || k->type == KSP_FGMRES || k->type == KSP_MINRES;
    if (!nt.empty() && nt != "default" && !(nt == natural) && !(nt == "none" && !krylov))
      throw Error(B200SP_ERR_UNSUPPORTED, "-" + prefix + "ksp_norm_type " + nt + ": " + tname + " monitors the " + natural + " residual norm only");
  }
  if (k->type == KSP_CHEBYSHEV) {
    const std::string ev = opt(prefix + "ksp_chebyshev_eigenvalues", "");
    if (!ev.empty()) {
      const size_t comma = ev.find(',');
      B2_REQUIRE(comma != std::string::npos, "-ksp_chebyshev_eigenvalues needs emin,emax");
      k->emin = std::stod(ev.substr(0, comma));
      k->emax = std::stod(ev.substr(comma + 1));
    } else {
      const double lam = estimate_lambda_max(ctx, A, M, 10);
      k->emin = 0.1 * lam;
      k->emax = 1.1 * lam;
    }
  }
  return k;
}

Op *Solver::make_simple_pc(const std::string &prefix, std::shared_ptr<Csr> mat, const char *default_type) {
  // PETSc's default PC on an assembled AIJ matrix is ILU(0) (block Jacobi + ILU(0) on more than one rank), which is not
  // on this library's kernel list: an unspecified -pc_type is an error, not a silent substitute
  if (!has(prefix + "pc_type") && std::string(default_type) == "petsc-default")
    throw Error(B200SP_ERR_UNSUPPORTED, "-" + prefix + "pc_type not given: PETSc would use ILU(0) here, which this library does not provide; "
                                        "choose one of none, jacobi, mg, gamg, lu" + (prefix.empty() ? ", fieldsplit" : ""));
  const std::string t = opt(prefix + "pc_type", default_type);
  if (t == "none") return nullptr;
  if (t == "jacobi") return add_op<JacobiOp>(*mat);
  if (t == "lu") return add_op<DenseInvOp>(*mat);
  if (t == "mg") return make_mg(prefix, mat);
  if (t == "gamg") return make_gamg(prefix, mat);
  throw Error(B200SP_ERR_UNSUPPORTED, "unsupported -" + prefix + "pc_type " + t);
}

// PCMG: rediscretised coarse velocity operators (the same device assembly on the coarser DMDA + the same
// Dirichlet elimination), Q1 interpolation with Dirichlet rows/cols zeroed, R = P^T, Chebyshev/Jacobi smoothing.
// Builds levels [first .. first+nlev-1] of a single-rank (or replicated) hierarchy into mg; A0 may be null
// (then the first level is assembled too).  local_only: the data is replicated on every rank, so the eigenvalue
// estimates must not be all-reduced.
static double smoother_setup(Solver *S, Ksp *k, const std::string &sp, Op *Aop, Op *jac, bool local_only,
                           const std::map<std::string, std::string> &opts) {
  auto opt = [&](const std::string &key, const std::string &def) { S->used[key] = true; auto it = opts.find(key); return it == opts.end() ? def : it->second; };
  // PETSc's PCMG default smoother is Chebyshev + SOR; SOR is not on this library's kernel list, so the level PC is
  // Jacobi and any other explicit choice is rejected
  const std::string lpc = opt(sp + "pc_type", "jacobi");
  if (lpc != "jacobi") throw Error(B200SP_ERR_UNSUPPORTED, "-" + sp + "pc_type " + lpc + ": the multigrid smoother preconditioner is jacobi only");
  k->set_operators(Aop, jac);
  const std::string t = opt(sp + "ksp_type", "chebyshev");
  k->type = t == "richardson" ? KSP_RICHARDSON : KSP_CHEBYSHEV;
  B2_REQUIRE(t == "chebyshev" || t == "richardson", "mg smoother: -mg_levels_ksp_type must be chebyshev or richardson");
  k->max_it = std::stoi(opt(sp + "ksp_max_it", "2"));
  k->norm_none = true;
  k->richardson_scale = std::stod(opt(sp + "ksp_richardson_scale", "1.0"));
  const double lam = estimate_lambda_max(S->ctx, Aop, jac, 10, local_only);
  k->emin = 0.1 * lam;
  k->emax = 1.1 * lam;
  return lam;
}

void Solver::build_levels_single(MgOp *mg, std::shared_ptr<Csr> A0, int Ml, int Nl, int nlev, const std::string &prefix, bool local_only) {
  const std::string sp = prefix + "mg_levels_";
  std::shared_ptr<Csr> Al = A0;
  if (!Al) {
    Dmda d0;
    d0.ctx = ctx; d0.M = Ml; d0.N = Nl; d0.xm = Ml; d0.ym = Nl;
    Al = assemble_stress(d0, 0);
    std::vector<int> ids = dmda_bc_ids(d0, 2);
    csr_zero_rows_cols(*Al, (int)ids.size(), ids.data(), 1.0, true, true, true);
    Al->tag = "spmv:A_coarse";
  }
  for (int l = 0; l < nlev; ++l) {
    auto L = std::make_unique<MgOp::Level>();
    L->A = Al;
    L->b.alloc((size_t)Al->nrows + 2); L->x.alloc((size_t)Al->nrows + 2); L->r.alloc((size_t)Al->nrows + 2);
    if (l < nlev - 1) {
      B2_REQUIRE((Ml - 1) % 2 == 0 && (Nl - 1) % 2 == 0 && Ml >= 5 && Nl >= 5, "pc mg: grid not coarsenable to the requested number of levels");
      const int Mc = (Ml - 1) / 2 + 1, Nc = (Nl - 1) / 2 + 1;
      L->P = interp_q1(ctx, Mc, Nc, 2, 1);
      L->R = restrict_q1(ctx, Mc, Nc, 2, 1);
      L->jac = std::make_unique<JacobiOp>(*Al);
      L->Aop = std::make_unique<CsrOp>(Al);
      L->smooth = std::make_unique<Ksp>(ctx, sp);
      smoother_setup(this, L->smooth.get(), sp, L->Aop.get(), L->jac.get(), local_only, opts);
      Dmda dc;
      dc.ctx = ctx; dc.M = Mc; dc.N = Nc; dc.xm = Mc; dc.ym = Nc;
      auto Ac = assemble_stress(dc, 0);
      std::vector<int> ids = dmda_bc_ids(dc, 2);
      csr_zero_rows_cols(*Ac, (int)ids.size(), ids.data(), 1.0, true, true, true);
      Ac->tag = "spmv:A_coarse";
      Al = Ac;
      Ml = Mc; Nl = Nc;
    }
    mg->lev.push_back(std::move(L));
  }
  mg->coarse = std::make_unique<DenseInvOp>(*mg->lev.back()->A);
}

Op *Solver::make_mg(const std::string &prefix, std::shared_ptr<Csr> mat) {
  if (mat->grid_M == 0 && have_grid && !mat->halo && (int64_t)2 * grid_M * grid_N == mat->nrows && mat->nrows == mat->ncols) {
    mat->grid_M = grid_M; mat->grid_N = grid_N; mat->dof_r = mat->dof_c = 2; // KSPSetDM equivalent (b200sp_ksp_set_dmda)
  }
  B2_REQUIRE(mat->grid_M > 0 && mat->dof_r == 2 && mat->dof_c == 2, "pc mg: needs the velocity block on a DMDA (b200sp_assemble_stress, b200sp_mat_set_grid or b200sp_ksp_set_dmda)");
  const int nlev = std::stoi(opt(prefix + "pc_mg_levels", "2"));
  B2_REQUIRE(nlev >= 2, "pc mg: need at least 2 levels");
  MgOp *mg = add_op<MgOp>(ctx, (int64_t)mat->nrows);
  if (!ctx->dcomm) {
    build_levels_single(mg, mat, mat->grid_M, mat->grid_N, nlev, prefix, false);
    return mg;
  }
  // ---- row-partitioned hierarchy: the top `nd` levels are distributed (halo exchanges), everything below is
  // gathered and solved redundantly on every rank (no communication on the small levels; the coarse work is
  // replicated instead of scattered across ranks that would each own a handful of nodes)
  B2_REQUIRE(mat->layout && mat->halo, "pc mg: distributed matrix without a DMDA layout");
  const std::string sp = prefix + "mg_levels_";
  int nd = std::min(std::stoi(opt(prefix + "pc_mg_distributed_levels", "3")), nlev - 1);
  B2_REQUIRE(nd >= 1, "pc mg: need at least one distributed level");
  std::vector<std::shared_ptr<Layout>> lay{mat->layout};
  for (int l = 1; l <= nd; ++l) {
    try { lay.push_back(std::make_shared<Layout>(lay.back()->coarsen())); }
    catch (const Error &) { nd = l - 1; break; }
  }
  B2_REQUIRE(nd >= 1, "pc mg: the grid cannot be coarsened on this process grid (every rank must own >= 2 coarse nodes per direction)");
  std::vector<Dmda> d((size_t)nd + 1);
  for (int l = 0; l <= nd; ++l) {
    Dmda &q = d[(size_t)l];
    q.ctx = ctx; q.M = lay[(size_t)l]->M; q.N = lay[(size_t)l]->N; q.pm = lay[(size_t)l]->m; q.pn = lay[(size_t)l]->n;
    lay[(size_t)l]->box(ctx->rank, &q.xs, &q.ys, &q.xm, &q.ym);
    q.layout = lay[(size_t)l];
    q.halo = l == 0 ? mat->halo : make_halo(ctx, *lay[(size_t)l], ctx->rank);
  }
  std::shared_ptr<Csr> Al = mat;
  for (int l = 0; l < nd; ++l) {
    auto L = std::make_unique<MgOp::Level>();
    L->A = Al;
    L->b.alloc((size_t)Al->nrows + 2); L->x.alloc((size_t)Al->nrows + 2); L->r.alloc((size_t)Al->nrows + 2);
    L->P = interp_q1_dist(d[(size_t)l], d[(size_t)l + 1], 2, 1);
    L->R = restrict_q1_dist(d[(size_t)l], d[(size_t)l + 1], 2, 1);
    L->jac = std::make_unique<JacobiOp>(*Al);
    L->Aop = std::make_unique<CsrOp>(Al);
    L->smooth = std::make_unique<Ksp>(ctx, sp);
    smoother_setup(this, L->smooth.get(), sp, L->Aop.get(), L->jac.get(), false, opts);
    if (l + 1 < nd) {
      auto Ac = assemble_stress(d[(size_t)l + 1], 0);
      std::vector<int> ids = dmda_bc_ids(d[(size_t)l + 1], 2);
      csr_zero_rows_cols(*Ac, (int)ids.size(), ids.data(), 1.0, true, true, true);
      Ac->tag = "spmv:A_coarse";
      Al = Ac;
    }
    mg->lev.push_back(std::move(L));
  }
  // replicated part: levels nd .. nlev-1 on the natural-ordered global coarse grid
  const Layout &Lc = *lay[(size_t)nd];
  mg->replicated = std::make_unique<MgOp>(ctx, (int64_t)2 * Lc.M * Lc.N);
  build_levels_single(mg->replicated.get(), nullptr, Lc.M, Lc.N, nlev - nd, prefix, true);
  // bridge maps: gathered (rank-major, padded to the largest rank) -> natural ordering, and natural -> my owned part
  int cnt_max = 0;
  for (int q = 0; q < Lc.size; ++q) cnt_max = std::max(cnt_max, 2 * Lc.lx[(size_t)(q % Lc.m)] * Lc.ly[(size_t)(q / Lc.m)]);
  std::vector<int> gmap((size_t)cnt_max * Lc.size, -1), lmap;
  for (int q = 0; q < Lc.size; ++q) {
    int xs, ys, xm, ym;
    Lc.box(q, &xs, &ys, &xm, &ym);
    for (int j = 0; j < ym; ++j)
      for (int i = 0; i < xm; ++i)
        for (int c = 0; c < 2; ++c) {
          const int nat = ((ys + j) * Lc.M + xs + i) * 2 + c;
          gmap[(size_t)q * cnt_max + (size_t)(j * xm + i) * 2 + c] = nat;
          if (q == ctx->rank) lmap.push_back(nat);
        }
  }
  mg->bridge_cnt = cnt_max;
  mg->bridge.reset(make_collective(ctx, cnt_max));
  mg->bridge_nloc = (int)lmap.size();
  mg->gather_map.alloc(gmap.size() + 1);
  mg->local_map.alloc(lmap.size() + 1);
  B2_CUDA(cudaMemcpyAsync(mg->gather_map.p, gmap.data(), sizeof(int) * gmap.size(), cudaMemcpyHostToDevice, ctx->stream));
  B2_CUDA(cudaMemcpyAsync(mg->local_map.p, lmap.data(), sizeof(int) * lmap.size(), cudaMemcpyHostToDevice, ctx->stream));
  mg->loc_b.alloc((size_t)cnt_max + 2); mg->loc_x.alloc((size_t)cnt_max + 2);
  mg->loc_b.zero(ctx->stream);
  mg->g_all.alloc((size_t)cnt_max * Lc.size + 2);
  mg->nat_b.alloc((size_t)2 * Lc.M * Lc.N + 2); mg->nat_x.alloc((size_t)2 * Lc.M * Lc.N + 2);
  ctx->sync();
  return mg;
}

// PCGAMG analogue: smoothed-aggregation hierarchy built on the device (kernels_amg.cu) from the assembled block alone
// -- no grid needed, so it also serves matrices that did not come from a DMDA (selfp / LSC products, user CSR).
// Options (PETSc's names): -pc_gamg_threshold (0), -pc_gamg_agg_nsmooths (1), -pc_gamg_coarse_eq_limit (50),
// -pc_mg_levels (30, the maximum), smoothers under -mg_levels_; ours: -pc_gamg_block_size (the matrix block size),
// -pc_gamg_mis_ordering {hash,natural} (PETSc's greedy MIS uses a random permutation = hash).
Op *Solver::make_gamg(const std::string &prefix, std::shared_ptr<Csr> mat) {
  if (mat->halo || ctx->dcomm) throw Error(B200SP_ERR_UNSUPPORTED, "-" + prefix + "pc_type gamg: the aggregation set-up is single-rank; use -" + prefix + "pc_type mg on row-partitioned DMDA matrices");
  B2_REQUIRE(mat->nrows == mat->ncols, "pc gamg: square matrix expected");
  const int bs0 = mat->dof_r > 0 && mat->dof_r == mat->dof_c ? mat->dof_r : 1;
  const int bs = std::stoi(opt(prefix + "pc_gamg_block_size", std::to_string(bs0)));
  const double theta = std::stod(opt(prefix + "pc_gamg_threshold", "0"));
  const int nsmooths = std::stoi(opt(prefix + "pc_gamg_agg_nsmooths", "1"));
  const std::string ord = opt(prefix + "pc_gamg_mis_ordering", "hash");
  B2_REQUIRE(ord == "hash" || ord == "natural", "pc gamg: -pc_gamg_mis_ordering must be hash or natural");
  const int order = ord == "natural" ? 1 : 0;
  const int coarse_limit = std::stoi(opt(prefix + "pc_gamg_coarse_eq_limit", "50"));
  const int max_levels = std::stoi(opt(prefix + "pc_mg_levels", "30"));
  B2_REQUIRE(nsmooths == 0 || nsmooths == 1, "pc gamg: -pc_gamg_agg_nsmooths must be 0 or 1");
  B2_REQUIRE(max_levels >= 1 && coarse_limit >= 1, "pc gamg: bad -pc_mg_levels / -pc_gamg_coarse_eq_limit");
  const std::string sp = prefix + "mg_levels_";
  MgOp *mg = add_op<MgOp>(ctx, (int64_t)mat->nrows);
  mg->kind = "smoothed aggregation, Galerkin coarse operators";
  std::shared_ptr<Csr> Al = mat;
  DevBuf<int> w, wc; // finest-level nodes behind every node of the current / next level (empty: ones)
  for (;;) {
    auto L = std::make_unique<MgOp::Level>();
    L->A = Al;
    L->b.alloc((size_t)Al->nrows + 2); L->x.alloc((size_t)Al->nrows + 2); L->r.alloc((size_t)Al->nrows + 2);
    bool coarsen = (int)mg->lev.size() + 1 < max_levels && Al->nrows > coarse_limit;
    DevBuf<int> agg;
    int nagg = 0;
    if (coarsen) {
      nagg = amg_aggregate(*Al, bs, theta, order, agg);
      if (nagg == 0 || (int64_t)nagg * bs >= Al->nrows) coarsen = false; // nothing left to aggregate: this level is the coarse one
    }
    if (!coarsen) { mg->lev.push_back(std::move(L)); break; }
    L->jac = std::make_unique<JacobiOp>(*Al);
    L->Aop = std::make_unique<CsrOp>(Al);
    L->smooth = std::make_unique<Ksp>(ctx, sp);
    const double lam = smoother_setup(this, L->smooth.get(), sp, L->Aop.get(), L->jac.get(), false, opts);
    auto Pt = amg_tentative(ctx, Al->nrows / bs, bs, agg, nagg, w.p, wc);
    w = std::move(wc);
    // omega = 4 / (3 lambda), lambda = the smoother's estimate of lambda_max(D^-1 A)
    L->P = nsmooths ? amg_smooth_prolongator(*Al, *Pt, 4.0 / (3.0 * lam)) : Pt;
    L->R = csr_transpose(*L->P);
    L->P->tag = "spmv:P"; L->R->tag = "spmv:R";
    auto AP = csr_matmat(*Al, *L->P);
    auto Ac = csr_matmat(*L->R, *AP);
    Ac->tag = "spmv:A_coarse";
    mg->lev.push_back(std::move(L));
    Al = Ac;
  }
  B2_REQUIRE(mg->lev.back()->A->nrows <= 8192, "pc gamg: the coarsest level has " + std::to_string(mg->lev.back()->A->nrows) +
             " rows, too many for the dense coarse solve; raise -" + prefix + "pc_mg_levels or lower -" + prefix + "pc_gamg_threshold");
  mg->coarse = std::make_unique<DenseInvOp>(*mg->lev.back()->A);
  return mg;
}

Op *Solver::make_fieldsplit() {
  B2_REQUIRE(opt("pc_fieldsplit_type", "schur") == "schur", "pc fieldsplit: only -pc_fieldsplit_type schur");
  std::shared_ptr<Csr> A00, A01, A10, A11;
  std::vector<int> strided_map; // non-empty: monolithic matrix, splits defined by strided fields
  if (Pmat->nest) {
    A00 = Pmat->blk[0][0]; A01 = Pmat->blk[0][1]; A10 = Pmat->blk[1][0]; A11 = Pmat->blk[1][1];
  } else {
    // the reference's case: KSPSetOperators(A, A) on the DMDA matrix (block size 2, no DM on the KSP), so PCFIELDSPLIT
    // defines the splits from the block size: field k -> split k, or -pc_fieldsplit_block_size / -pc_fieldsplit_1_fields
    auto P = Pmat->csr;
    const int bs = std::stoi(opt("pc_fieldsplit_block_size", std::to_string(P->dof_r > 0 ? P->dof_r : 2)));
    B2_REQUIRE(bs >= 2 && bs <= 8 && P->nrows == P->ncols && P->nrows % bs == 0, "pc fieldsplit: bad -pc_fieldsplit_block_size for this matrix");
    std::vector<int> split((size_t)bs, 0);
    std::string f1 = opt("pc_fieldsplit_1_fields", std::to_string(bs - 1));
    for (char &ch : f1) if (ch == ',') ch = ' ';
    std::istringstream in(f1);
    int f;
    while (in >> f) { B2_REQUIRE(f >= 0 && f < bs, "pc fieldsplit: field out of range"); split[(size_t)f] = 1; }
    if (!has("pc_fieldsplit_1_fields") && bs > 2) throw Error(B200SP_ERR_UNSUPPORTED, "pc fieldsplit: block size > 2 needs -pc_fieldsplit_1_fields");
    A00 = csr_extract_fields(*P, bs, split, 0, 0); A01 = csr_extract_fields(*P, bs, split, 0, 1);
    A10 = csr_extract_fields(*P, bs, split, 1, 0); A11 = csr_extract_fields(*P, bs, split, 1, 1);
    A00->tag = "spmv:A00"; A01->tag = "spmv:A01"; A10->tag = "spmv:A10"; A11->tag = "spmv:A11";
    mats.push_back(A00); mats.push_back(A01); mats.push_back(A10); mats.push_back(A11);
    const int nnode = P->nrows / bs;
    int nf0 = 0;
    for (int k = 0; k < bs; ++k) nf0 += split[(size_t)k] == 0;
    strided_map.resize((size_t)P->nrows);
    for (int node = 0; node < nnode; ++node) {
      int p0 = 0, p1 = 0;
      for (int k = 0; k < bs; ++k) {
        if (split[(size_t)k] == 0) strided_map[(size_t)node * nf0 + p0++] = node * bs + k;
        else strided_map[(size_t)nnode * nf0 + (size_t)node * (bs - nf0) + p1++] = node * bs + k;
      }
    }
  }
  const std::string fs = opt("pc_fieldsplit_schur_fact_type", "full");
  const int fact = fs == "diag" ? 0 : fs == "lower" ? 1 : fs == "upper" ? 2 : fs == "full" ? 3 : -1;
  B2_REQUIRE(fact >= 0, "bad -pc_fieldsplit_schur_fact_type " + fs);
  const std::string pre = opt("pc_fieldsplit_schur_precondition", "a11");
  const double scale = std::stod(opt("pc_fieldsplit_schur_scale", "-1.0"));
  // K0: -fieldsplit_0_ KSP on A00
  Op *A00op = add_op<CsrOp>(A00);
  Op *pc0 = make_simple_pc("fieldsplit_0_", A00, "petsc-default");
  Ksp *k0 = make_ksp("fieldsplit_0_", A00op, pc0, "preonly");
  Op *K0 = add_op<KspOp>(k0);
  // S with its own identically configured inner KSP (MatSchurComplementGetKSP)
  Ksp *k0s = make_ksp("fieldsplit_0_", A00op, pc0, "preonly");
  Op *K0s = add_op<KspOp>(k0s);
  Op *S = add_op<SchurOp>(A11, A10, K0s, A01);
  // matrix the S-solve's PC is built from
  std::shared_ptr<Csr> Sp;
  if (pre == "a11") { B2_REQUIRE(A11 != nullptr, "schur precondition a11: the nest has no (1,1) block"); Sp = A11; }
  else if (pre == "user") { B2_REQUIRE(schur_user != nullptr, "schur precondition user: call b200sp_ksp_set_schur_user_mat"); Sp = schur_user; }
  else if (pre == "selfp") { // Sp = A11 - A10 diag(A00)^-1 A01
    JacobiOp dj(*A00);
    auto A10D = csr_scale_cols(*A10, dj.dinv.p);
    auto prod = csr_matmat(*A10D, *A01);
    // no (1,1) block (the reference's [A Bt; B 0]): Sp = -A10 D^-1 A01 = prod + (-2) prod, exact in floating point
    Sp = A11 ? csr_add_scaled(*A11, -1.0, *prod) : csr_add_scaled(*prod, -2.0, *prod);
    mats.push_back(Sp);
  } else if (pre != "self") throw Error(B200SP_ERR_UNSUPPORTED, "unsupported -pc_fieldsplit_schur_precondition " + pre);
  const std::string pt = opt("fieldsplit_1_pc_type", Sp ? "petsc-default" : "none");
  Op *pcS = nullptr;
  if (pt == "lsc") {
    const bool sd = has("fieldsplit_1_pc_lsc_scale_diag");
    std::shared_ptr<Csr> Lm;
    if (sd) {
      JacobiOp dj(*A00);
      auto A10D = csr_scale_cols(*A10, dj.dinv.p);
      Lm = csr_matmat(*A10D, *A01);
    } else Lm = csr_matmat(*A10, *A01);
    mats.push_back(Lm);
    Op *Lop = add_op<CsrOp>(Lm);
    Op *pcl = make_simple_pc("fieldsplit_1_lsc_", Lm, "petsc-default");
    Ksp *kl = make_ksp("fieldsplit_1_lsc_", Lop, pcl, "gmres"); // PCLSC creates a fresh KSP: GMRES unless told otherwise
    Op *Linv = add_op<KspOp>(kl);
    pcS = add_op<LscOp>(A00, A01, A10, Linv, sd);
  } else if (pt != "none" && Sp) {
    pcS = make_simple_pc("fieldsplit_1_", Sp, "petsc-default");
  }
  Ksp *kS = make_ksp("fieldsplit_1_", S, pcS, "gmres"); // the Schur KSP is a fresh KSP in PETSc: GMRES unless told otherwise
  Op *KS = add_op<KspOp>(kS);
  Op *fsop = add_op<FieldSplitOp>(fact, scale, A01, A10, K0, KS);
  if (!strided_map.empty()) return add_op<StridedSplitOp>(fsop, strided_map);
  return fsop;
}

void Solver::setup() {
  B2_REQUIRE(Amat && Pmat, "KSPSetUp: operators not set");
  outer = nullptr; outer_pc = nullptr; is_setup = false; // a setup() that throws must not leave pointers into the cleared trees
  ops.clear(); ksps.clear(); mats.clear();
  used.clear();
  Op *Aop = Amat->nest ? (Op *)add_op<NestOp>(Amat->blk[0][0], Amat->blk[0][1], Amat->blk[1][0], Amat->blk[1][1]) : (Op *)add_op<CsrOp>(Amat->csr);
  const std::string pt = opt("pc_type", "petsc-default");
  if (pt == "petsc-default")
    throw Error(B200SP_ERR_UNSUPPORTED, "-pc_type not given: PETSc would use ILU(0) (block Jacobi + ILU(0) in parallel), which this library does not "
                                        "provide; choose one of none, jacobi, mg, lu, fieldsplit");
  Op *pc = nullptr;
  if (pt == "fieldsplit") {
    B2_REQUIRE(Pmat->nest || !ctx->dcomm, "pc fieldsplit on a monolithic matrix: single GPU only (use the nest layout when row-partitioned)");
    pc = make_fieldsplit();
  } else {
    B2_REQUIRE(!Pmat->nest, "pc " + pt + " on a nest matrix: use -pc_type fieldsplit");
    pc = make_simple_pc("", Pmat->csr, "petsc-default");
  }
  outer_pc = pc;
  outer = make_ksp("", Aop, pc, "gmres");
  outer->keep_history = true;
  outer->use_pc_graph = pc && opt("b200sp_pc_graph", "1") != "0" && pc->capturable() && (!ctx->dcomm || ctx->dcomm->capturable());
  check_options_left();
  ctx->sync();
  setup_state = Amat->state() + Pmat->state() + (schur_user ? schur_user->state : 0);
  is_setup = true;
}

std::string Solver::view() const { return outer ? outer->view(0) : std::string("KSP not set up\n"); }

} // namespace b200sp
