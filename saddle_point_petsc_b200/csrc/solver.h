// solver.h -- host-side orchestration objects: linear operators, preconditioners and Krylov drivers.
// These mirror the PETSc objects the reference reaches through KSPSetFromOptions/KSPSetUp/KSPSolve
// (src/SaddlePointProblem.c:65-72); all arithmetic is in the CUDA kernels they launch.
#pragma once
#include "core.h"
#include "dist.h"

namespace b200sp {

// y = Op(x) on device pointers (x and y never alias)
struct Op {
  Ctx *ctx;
  int64_t n_in, n_out;
  Op(Ctx *c, int64_t ni, int64_t no) : ctx(c), n_in(ni), n_out(no) {}
  virtual ~Op() {}
  virtual void apply(const double *x, double *y) = 0;
  // r = b - Op(x); default: apply then r = b - r
  virtual void residual(const double *b, const double *x, double *r) {
    apply(x, r);
    vec_aypx(ctx, n_out, -1.0, b, r);
  }
  virtual std::string view(int indent) const = 0;
  virtual const double *jacobi_dinv() const { return nullptr; } // non-null when the op is y = x .* dinv
  virtual const Csr *csr() const { return nullptr; }            // non-null when the op is a plain CSR MatMult
  // true when apply() only enqueues work on the context's streams (no host synchronisation, no allocation after
  // the first call), i.e. it can be recorded into a CUDA graph
  virtual bool capturable() const { return true; }
};

struct CsrOp : Op { // MatMult
  std::shared_ptr<Csr> A;
  explicit CsrOp(std::shared_ptr<Csr> a) : Op(a->ctx, a->ncols, a->nrows), A(a) {}
  void apply(const double *x, double *y) override { csr_spmv(*A, x, y); }
  void residual(const double *b, const double *x, double *r) override { csr_spmv(*A, x, r, -1.0, b, 1.0); }
  const Csr *csr() const override { return A.get(); }
  std::string view(int indent) const override;
};

struct NestOp : Op { // MATNEST 2x2 on [x0; x1]
  std::shared_ptr<Csr> b00, b01, b10, b11;
  NestOp(std::shared_ptr<Csr> a00, std::shared_ptr<Csr> a01, std::shared_ptr<Csr> a10, std::shared_ptr<Csr> a11)
      : Op(a00->ctx, a00->ncols + a01->ncols, a00->nrows + a10->nrows), b00(a00), b01(a01), b10(a10), b11(a11) {}
  void apply(const double *x, double *y) override;
  void residual(const double *b, const double *x, double *r) override;
  std::string view(int indent) const override;
};

struct JacobiOp : Op { // PCJACOBI
  DevBuf<double> dinv;
  explicit JacobiOp(const Csr &A);
  void apply(const double *x, double *y) override { vec_pointwise_mult(ctx, n_in, x, dinv.p, y); }
  std::string view(int indent) const override;
  const double *jacobi_dinv() const override { return dinv.p; }
};

struct DenseInvOp : Op { // exact coarse solve (PCLU stand-in): explicit inverse, dense mat-vec
  DevBuf<double> Ainv;
  explicit DenseInvOp(const Csr &A);
  void apply(const double *x, double *y) override { dense_matvec(ctx, (int)n_in, Ainv.p, x, y); }
  std::string view(int indent) const override;
};

struct Ksp;

struct KspOp : Op { // y = ksp(x), zero initial guess
  Ksp *ksp;
  explicit KspOp(Ksp *k);
  bool capturable() const override;
  void apply(const double *x, double *y) override;
  std::string view(int indent) const override;
};

struct SchurOp : Op { // MatSchurComplement: S x = A11 x - A10 ksp(A00) A01 x
  std::shared_ptr<Csr> A11, A10, A01;
  Op *K0;
  DevBuf<double> t0, t1;
  SchurOp(std::shared_ptr<Csr> a11, std::shared_ptr<Csr> a10, Op *k0, std::shared_ptr<Csr> a01);
  bool capturable() const override { return K0->capturable(); }
  void apply(const double *x, double *y) override;
  std::string view(int indent) const override;
};

struct FieldSplitOp : Op { // PCFIELDSPLIT, Schur factorisations
  int fact; // 0 diag, 1 lower, 2 upper, 3 full
  double scale;
  std::shared_ptr<Csr> A01, A10;
  Op *K0, *KS;
  DevBuf<double> t0, t1;
  FieldSplitOp(int fact_, double scale_, std::shared_ptr<Csr> a01, std::shared_ptr<Csr> a10, Op *k0, Op *ks);
  bool capturable() const override { return K0->capturable() && KS->capturable(); }
  void apply(const double *b, double *y) override;
  std::string view(int indent) const override;
};

struct StridedSplitOp : Op { // fieldsplit on a monolithic (field-interleaved) vector: gather the splits, apply, scatter back
  Op *inner;
  DevBuf<int> map;          // split-ordered position -> interleaved position
  DevBuf<double> xs, ys;
  StridedSplitOp(Op *in, const std::vector<int> &m);
  bool capturable() const override { return inner->capturable(); }
  void apply(const double *x, double *y) override;
  std::string view(int indent) const override;
};

struct LscOp : Op { // PCLSC
  std::shared_ptr<Csr> A00, A01, A10;
  Op *Linv;
  DevBuf<double> dinv; // empty unless scale_diag
  bool scale_diag;
  DevBuf<double> p0, p1, u0, u1;
  LscOp(std::shared_ptr<Csr> a00, std::shared_ptr<Csr> a01, std::shared_ptr<Csr> a10, Op *linv, bool scale_diag_);
  bool capturable() const override { return Linv->capturable(); }
  void apply(const double *x, double *y) override;
  std::string view(int indent) const override;
};

struct MgOp : Op { // PCMG multiplicative V-cycle
  struct Level {
    std::shared_ptr<Csr> A, P, R; // P: level l+1 -> l ; R = P^T
    std::unique_ptr<JacobiOp> jac;
    std::unique_ptr<CsrOp> Aop;
    std::unique_ptr<Ksp> smooth;
    DevBuf<double> b, x, r;
  };
  std::vector<std::unique_ptr<Level>> lev;
  std::unique_ptr<DenseInvOp> coarse;
  std::string kind = "rediscretised coarse operators, Q1 interpolation";
  // row-partitioned runs: `lev` holds the distributed smoothing levels; the levels below are replicated on every rank
  std::unique_ptr<MgOp> replicated;
  DevBuf<double> loc_b, loc_x, g_all, nat_b, nat_x;
  DevBuf<int> gather_map, local_map;
  int bridge_cnt = 0, bridge_nloc = 0;
  std::unique_ptr<struct Collective> bridge; // all-gather of the restricted residual (peer-to-peer when possible)
  MgOp(Ctx *c, int64_t n) : Op(c, n, n) {}
  bool capturable() const override;
  void cycle(int l, const double *b, double *x);
  void apply(const double *b, double *x) override;
  std::string view(int indent) const override;
};

enum GmresOrthog { ORTHOG_CGS = 0, ORTHOG_CGS_REFINE_IFNEEDED = 1, ORTHOG_CGS_REFINE_ALWAYS = 2, ORTHOG_MGS = 3 };
enum KspType { KSP_PREONLY = 0, KSP_RICHARDSON = 1, KSP_CHEBYSHEV = 2, KSP_GMRES = 3, KSP_FGMRES = 4, KSP_MINRES = 5 };

struct Ksp {
  Ctx *ctx;
  std::string prefix;
  int type = KSP_GMRES;
  Op *A = nullptr, *M = nullptr; // M == nullptr: identity
  int64_t n = 0;
  double rtol = 1e-5, atol = 1e-50, dtol = 1e5; // PETSc defaults
  int max_it = 10000, restart = 30;
  bool norm_none = false;
  double emin = 0, emax = 0, richardson_scale = 1.0;
  bool keep_history = false;
  int orthog = ORTHOG_CGS; // -ksp_gmres_modifiedgramschmidt / -ksp_gmres_cgs_refinement_type (PETSc default: CGS, refine_never)
  // results
  int its = 0, reason = 0;
  double rnorm = 0, rnorm0 = 0;
  std::vector<double> hist;
  // workspace (allocated on first solve)
  DevBuf<double> V, Z, w0, w1, w2, w3, w4, w5, w6, w7, w8;
  int64_t ld = 0;
  // CUDA-graph replay of the preconditioner application (outer KSP only): one instantiated graph per (input, output)
  // vector pair -- the Krylov basis vectors are persistent, so the pairs repeat from solve to solve
  struct PcGraph { cudaGraphExec_t exec = nullptr; int64_t launches = 0; };
  // smoother inside PCMG on a row-partitioned level: the halo (and dof) of the matrix that multiplies this solver's
  // OUTPUT next (the level operator for the residual, or the finer level's interpolation), so the last sweep can push
  // its boundary values from inside the SpMV kernel (csr_spmv_epi push_to)
  Halo *push_after = nullptr;
  int push_after_dof = 0;
  bool use_pc_graph = false, pc_warmed = false;
  std::map<std::pair<const double *, double *>, PcGraph> pc_graphs;
  ~Ksp();

  Ksp(Ctx *c, const std::string &pfx) : ctx(c), prefix(pfx) {}
  void set_operators(Op *a, Op *m) { A = a; M = m; n = a->n_in; }
  int solve(const double *b, double *x, bool guess_nonzero);
  std::string view(int indent) const;

private:
  void pc_apply(const double *x, double *y);
  int converged(int it, double rn);
  double norm2(const double *x);
  double dot(const double *x, const double *y);
  int solve_richardson(const double *b, double *x, bool guess_nonzero);
  int solve_chebyshev(const double *b, double *x, bool guess_nonzero);
  int solve_gmres(const double *b, double *x, bool guess_nonzero, bool flexible);
  int solve_minres(const double *b, double *x, bool guess_nonzero);
};

// deterministic lambda_max estimate of M^-1 A (10 power iterations from the hashed vector), same procedure as
// the oracle's or_estimate_lambda_max; chebyshev bounds are (0.1, 1.1) x estimate (PETSc's default transform)
double estimate_lambda_max(Ctx *c, Op *A, Op *M, int nits, bool local_only = false);

// the object behind b200sp_ksp: options + operators + the composed solver tree
struct Solver {
  Ctx *ctx;
  std::map<std::string, std::string> opts;
  // operators are held BY VALUE: a Mat is a handle on shared CSR blocks, so the caller may destroy its matrix handles
  // before KSPSolve / KSPDestroy (PETSc reference-counts the operators of KSPSetOperators)
  Mat Amat_, Pmat_;
  Mat *Amat = nullptr, *Pmat = nullptr; // point at the copies above once the operators are set
  void set_operators(const Mat &A, const Mat &P) { Amat_ = A; Pmat_ = P; Amat = &Amat_; Pmat = &Pmat_; is_setup = false; }
  std::shared_ptr<Csr> schur_user;
  bool have_grid = false;
  int grid_M = 0, grid_N = 0;
  std::vector<std::unique_ptr<Op>> ops;
  std::vector<std::unique_ptr<Ksp>> ksps;
  std::vector<std::shared_ptr<Csr>> mats; // matrices built during setup (Sp, L, coarse levels)
  Ksp *outer = nullptr;
  Op *outer_pc = nullptr;
  bool is_setup = false;
  int64_t setup_state = 0; // Amat/Pmat value state at setup(); KSPSolve sets up again when the matrices changed since
  bool current() const { return is_setup && Amat && Pmat && Amat->state() + Pmat->state() + (schur_user ? schur_user->state : 0) == setup_state; }
  DevBuf<double> host_b, host_x; // device staging for b200sp_ksp_solve_host
  explicit Solver(Ctx *c) : ctx(c) {}
  // every key a solver object reads is recorded; setup() then rejects solver options nobody consumed (PETSc's
  // -options_left, promoted to an error: a mistyped or unsupported option must not silently select another solver)
  mutable std::map<std::string, bool> used;
  void set_options(const char *text);
  void setup();
  std::string view() const;

private:
  std::string opt(const std::string &key, const std::string &def) const;
  bool has(const std::string &key) const { used[key] = true; return opts.count(key) != 0; }
  void check_options_left() const;
  template <class T, class... Args> T *add_op(Args &&...args) {
    ops.emplace_back(new T(std::forward<Args>(args)...));
    return static_cast<T *>(ops.back().get());
  }
  Ksp *make_ksp(const std::string &prefix, Op *A, Op *M, const char *default_type);
  Op *make_simple_pc(const std::string &prefix, std::shared_ptr<Csr> mat, const char *default_type);
  Op *make_mg(const std::string &prefix, std::shared_ptr<Csr> mat);
  Op *make_gamg(const std::string &prefix, std::shared_ptr<Csr> mat); // aggregation multigrid (kernels_amg.cu)

public:
  void build_levels_single(MgOp *mg, std::shared_ptr<Csr> A0, int Ml, int Nl, int nlev, const std::string &prefix, bool local_only);

private:
  Op *make_fieldsplit();
};

} // namespace b200sp
