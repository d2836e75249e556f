"""CPU tests: the oracle's Krylov methods and preconditioners against their PUBLISHED MATHEMATICAL DEFINITIONS, evaluated
with dense numpy linear algebra that shares no code (and no recurrences) with the oracle.

PETSc is not installed (the Krylov/PC half of the oracle is "parity unpinned" against PETSc
itself), so this is the strongest independent anchor available: a Krylov method is pinned by WHAT it minimises over
WHICH subspace, whatever the implementation --
  KSPGMRES  (left PC):   x_k = argmin || M^-1 (b - A x) ||_2      over x_0 + K_k(M^-1 A, M^-1 r_0),  restarted
  KSPFGMRES (right PC):  x_k = argmin || b - A x ||_2             over x_0 + M^-1 K_k(A M^-1, r_0)   (constant linear M)
  KSPMINRES (SPD PC):    x_k = argmin || b - A x ||_{M^-1}        over x_0 + K_k(M^-1 A, M^-1 r_0)
  KSPCHEBYSHEV:          e_k = T_k((theta - B)/delta) / T_k(theta/delta) e_0,  B = M^-1 A
  PCFIELDSPLIT Schur:    the DIAG / LOWER / UPPER / FULL block factorisations of the PETSc manual
  PCMG V-cycle:          x <- S_post( x + P A_c^-1 R (b - A S_pre(b)) )
so the residual histories / results must agree with the dense evaluation to rounding.  No GPU."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import sp_oracle as so  # noqa: E402


def dense_op(op, n):
    """matrix of a linear oracle operator, column by column"""
    out = np.empty((n, n))
    e = np.zeros(n)
    y = np.empty(n)
    for j in range(n):
        e[j] = 1.0
        so.lib().or_op_apply(op, so.dptr(e), so.dptr(y))
        out[:, j] = y
        e[j] = 0.0
    return out


def arnoldi_min_residuals(B, r0, kmax):
    """min_y || r0 - B V_k y ||_2 for k = 1..kmax with V_k an orthonormal basis of K_k(B, r0): Arnoldi with
    re-orthogonalised modified Gram-Schmidt and a dense least-squares solve per k (no Givens recurrences).
    Returns the norms and a function giving the minimiser's Krylov update V_k y."""
    n = len(r0)
    beta = np.linalg.norm(r0)
    V = np.zeros((n, kmax + 1))
    H = np.zeros((kmax + 1, kmax))
    V[:, 0] = r0 / beta
    norms, ys = [], []
    for k in range(kmax):
        w = B @ V[:, k]
        for _ in range(2):
            for j in range(k + 1):
                h = V[:, j] @ w
                H[j, k] += h
                w = w - h * V[:, j]
        H[k + 1, k] = np.linalg.norm(w)
        rhs = np.zeros(k + 2)
        rhs[0] = beta
        y, *_ = np.linalg.lstsq(H[:k + 2, :k + 1], rhs, rcond=None)
        norms.append(np.linalg.norm(rhs - H[:k + 2, :k + 1] @ y))
        ys.append(y)
        if H[k + 1, k] < 1e-14 * beta:
            break
        V[:, k + 1] = w / H[k + 1, k]
    return np.array(norms), (lambda k: V[:, :k] @ ys[k - 1])


def velocity_problem():
    p = so.Problem(10, 8)
    return p, p.A.scipy().toarray(), p.f


@pytest.mark.parametrize("pc", ["none", "jacobi"])
def test_gmres_minimises_the_preconditioned_residual_with_restarts(pc):
    p, A, b = velocity_problem()
    n = len(b)
    restart = 12                                  # several restart cycles before convergence
    s = so.Solver(p, "-ksp_type gmres -ksp_gmres_restart %d -ksp_rtol 1e-9 -pc_type %s" % (restart, pc))
    r = s.solve()
    assert r["reason"] == 2
    Minv = np.eye(n) if pc == "none" else np.diag(1.0 / np.diag(A))
    B = Minv @ A
    # KSPGMRES logs the recomputed residual again at the start of every restart cycle (KSPGMRESCycle, it == 0), so the
    # history holds one extra entry per restart: its + 1 + (number of restarts) values
    x = np.zeros(n)
    ref, done = [], 0
    while done < r["its"]:
        r0 = Minv @ (b - A @ x)
        ref.append(np.linalg.norm(r0))
        k = min(restart, r["its"] - done)
        norms, upd = arnoldi_min_residuals(B, r0, k)
        ref += list(norms)
        x = x + upd(len(norms))
        done += len(norms)
    ref = np.array(ref)
    hist = r["history"]
    assert len(hist) == len(ref) == r["its"] + 1 + (r["its"] - 1) // restart
    assert np.max(np.abs(hist - ref) / ref) < 1e-6, np.max(np.abs(hist - ref) / ref)
    assert np.linalg.norm(r["x"] - x) < 1e-7 * np.linalg.norm(x)
    assert hist[-1] <= 1e-9 * hist[0] < hist[-2]          # stopped at the first iterate that meets rtol


def test_fgmres_with_a_fixed_pc_minimises_the_true_residual():
    p, A, b = velocity_problem()
    n = len(b)
    s = so.Solver(p, "-ksp_type fgmres -ksp_gmres_restart 30 -ksp_rtol 1e-9 -pc_type jacobi")
    r = s.solve()
    assert r["reason"] == 2 and r["its"] <= 30
    Minv = np.diag(1.0 / np.diag(A))
    norms, upd = arnoldi_min_residuals(A @ Minv, b.copy(), r["its"])
    ref = np.concatenate([[np.linalg.norm(b)], norms])
    hist = r["history"][:r["its"] + 1]
    assert np.max(np.abs(hist - ref) / ref) < 1e-6
    x_ref = Minv @ upd(r["its"])                             # x = M^-1 (V_k y)
    assert np.linalg.norm(r["x"] - x_ref) < 1e-7 * np.linalg.norm(x_ref)
    assert abs(np.linalg.norm(b - A @ r["x"]) - hist[-1]) < 1e-6 * hist[0]   # the reported norm IS the true residual


DIAG_JACOBI = ("-pc_type fieldsplit -pc_fieldsplit_type schur -pc_fieldsplit_schur_precondition user "
               "-fieldsplit_0_ksp_type preonly -fieldsplit_0_pc_type jacobi -fieldsplit_1_ksp_type preonly -fieldsplit_1_pc_type jacobi ")


def test_minres_minimises_the_residual_in_the_pc_norm():
    p = so.Problem(8, 8, kkt=True, rhs_kind=1)
    K = p.scipy_K().toarray()
    b = p.rhs
    n = len(b)
    assert np.abs(K - K.T).max() < 1e-13
    s = so.Solver(p, "-ksp_type minres -ksp_rtol 1e-8 -pc_fieldsplit_schur_fact_type diag " + DIAG_JACOBI)
    Minv = dense_op(s.ksp.contents.M, n)
    assert np.abs(Minv - Minv.T).max() < 1e-13 and np.linalg.eigvalsh(0.5 * (Minv + Minv.T)).min() > 0   # SPD, as MINRES needs
    r = s.solve()
    assert r["reason"] == 2
    L = np.linalg.cholesky(0.5 * (Minv + Minv.T))           # M^-1 = L L^T ;  ||r||_{M^-1} = ||L^T r||_2
    Kt, bt = L.T @ K @ L, L.T @ b                           # x = L xt
    norms, _ = arnoldi_min_residuals(Kt, bt, r["its"])
    ref = np.concatenate([[np.linalg.norm(bt)], norms])
    # KSPMINRES (PETSc <= 3.18, minres.c) starts its monitored norm at ||M^-1 r_0||_2 and multiplies it by |s_k| every
    # step; the quantity that really shrinks by |s_k| is ||r_k||_{M^-1}.  So the history is the minimal M^-1-norm
    # residual times the constant ||M^-1 r_0||_2 / ||r_0||_{M^-1} (the relative convergence test does not see it).
    quirk = np.linalg.norm(Minv @ b) / np.sqrt(b @ Minv @ b)
    hist = r["history"][:r["its"] + 1]
    tol = 1e-5 * ref * quirk + 1e-12 * hist[0]
    assert np.all(np.abs(hist - quirk * ref) < tol), np.max(np.abs(hist - quirk * ref) / (quirk * ref))
    res = b - K @ r["x"]
    assert abs(quirk * np.sqrt(abs(res @ Minv @ res)) - hist[-1]) < 1e-9 * hist[0]


@pytest.mark.parametrize("k", [1, 2, 3, 5, 8])
def test_chebyshev_iterate_is_the_scaled_chebyshev_polynomial(k):
    p, A, b = velocity_problem()
    d = np.diag(A)
    S = A / np.sqrt(np.outer(d, d))                         # D^-1/2 A D^-1/2, symmetric: same spectrum as D^-1 A
    lam, V = np.linalg.eigh(S)
    emin, emax = 0.1 * lam.max(), 1.1 * lam.max()           # PETSc's default transform of the estimate
    s = so.Solver(p, "-ksp_type chebyshev -ksp_max_it %d -ksp_norm_type none -ksp_chebyshev_eigenvalues %.17g,%.17g -pc_type jacobi" % (k, emin, emax))
    r = s.solve()
    theta, delta = 0.5 * (emax + emin), 0.5 * (emax - emin)

    def T(j, z):                                            # Chebyshev polynomial of the first kind, any real z
        z = np.asarray(z, dtype=float)
        out = np.empty_like(z)
        inside = np.abs(z) <= 1
        out[inside] = np.cos(j * np.arccos(z[inside]))
        zz = z[~inside]
        out[~inside] = np.sign(zz) ** j * np.cosh(j * np.arccosh(np.abs(zz)))
        return out

    pk = T(k, (theta - lam) / delta) / T(k, np.array([theta / delta]))[0]   # error polynomial on the spectrum
    x_exact = np.linalg.solve(A, b)
    e0 = np.sqrt(d) * x_exact                               # error of the zero guess in the symmetrised variables
    ek = V @ (pk * (V.T @ e0))
    x_ref = x_exact - ek / np.sqrt(d)
    assert np.linalg.norm(r["x"] - x_ref) < 1e-10 * np.linalg.norm(x_exact), k


@pytest.mark.parametrize("fact", ["diag", "lower", "upper", "full"])
def test_fieldsplit_schur_factorisations_match_the_block_formulas(fact):
    p = so.Problem(6, 5, kkt=True, rhs_kind=1)
    s = so.Solver(p, "-ksp_type gmres -pc_fieldsplit_schur_fact_type %s " % fact + DIAG_JACOBI)
    n0, n1 = p.nu, p.np_
    A, Bt, B = p.A.scipy().toarray(), p.Bt.scipy().toarray(), p.B.scipy().toarray()
    dq = p.Q.scipy().diagonal()
    Ainv = np.diag(1.0 / np.diag(A))                         # preonly + jacobi on A00
    Sinv = np.diag(1.0 / np.where(dq == 0, 1.0, dq))         # preonly + jacobi built from the user matrix
    rng = np.random.default_rng(3)
    b = rng.standard_normal(n0 + n1)
    b0, b1 = b[:n0], b[n0:]
    if fact == "diag":                                       # PCFieldSplitSetSchurScale default -1
        y0, y1 = Ainv @ b0, -1.0 * (Sinv @ b1)
    elif fact == "lower":
        y0 = Ainv @ b0
        y1 = Sinv @ (b1 - B @ y0)
    elif fact == "upper":
        y1 = Sinv @ b1
        y0 = Ainv @ (b0 - Bt @ y1)
    else:
        y0 = Ainv @ b0
        y1 = Sinv @ (b1 - B @ y0)
        y0 = Ainv @ (b0 - Bt @ y1)
    y = np.empty(n0 + n1)
    so.lib().or_op_apply(s.ksp.contents.M, so.dptr(b), so.dptr(y))
    ref = np.concatenate([y0, y1])
    assert np.max(np.abs(y - ref)) < 1e-12 * np.max(np.abs(ref)), fact


def test_two_level_mg_vcycle_matches_the_dense_error_propagation():
    p = so.Problem(8, 6)
    s = so.Solver(p, "-ksp_type fgmres -pc_type mg -pc_mg_levels 2")
    A = p.A.scipy().toarray()
    n = A.shape[0]
    Mc, Nc = 5, 4
    P = so.Csr(so.lib().or_interp_q1(Mc, Nc, 2, 1)).scipy().toarray()
    Ac = so.Csr(so.lib().or_assemble_A(Mc, Nc, 0))
    ids = so.bc_ids(Mc, Nc, 2)
    so.lib().or_apply_bc(Ac.ptr, None, len(ids), so.iptr(ids))
    Ac = Ac.scipy().toarray()
    sm = s.mg_smooth[0].contents
    emin, emax, k = sm.emin, sm.emax, sm.max_it
    assert k == 2
    # Chebyshev(k)/Jacobi smoother as an error-propagation matrix (definition checked in the test above)
    d = np.diag(A)
    lam, V = np.linalg.eigh(A / np.sqrt(np.outer(d, d)))
    assert 0.75 * lam.max() < emax / 1.1 <= lam.max() * (1 + 1e-12)   # ten power iterations: a lower estimate of lambda_max
    assert abs(emin / emax - 0.1 / 1.1) < 1e-15                        # PETSc's default transform (0, 0.1; 0, 1.1)
    theta, delta = 0.5 * (emax + emin), 0.5 * (emax - emin)
    z = (theta - lam) / delta
    tz = np.where(np.abs(z) <= 1, np.cos(k * np.arccos(np.clip(z, -1, 1))), np.sign(z) ** k * np.cosh(k * np.arccosh(np.maximum(np.abs(z), 1))))
    t0 = np.cosh(k * np.arccosh(theta / delta))
    Dh = np.diag(np.sqrt(d))
    Dhi = np.diag(1.0 / np.sqrt(d))
    Es = Dhi @ V @ np.diag(tz / t0) @ V.T @ Dh              # e <- Es e for one smoother call
    Ainv = np.linalg.inv(A)
    CGC = np.eye(n) - P @ np.linalg.solve(Ac, P.T @ A)      # coarse-grid correction, R = P^T
    E = Es @ CGC @ Es
    rng = np.random.default_rng(5)
    b = rng.standard_normal(n)
    b[p.bc] = 0.0                                           # Dirichlet rows are decoupled identities on every level
    y = np.empty(n)
    so.lib().or_op_apply(s.ksp.contents.M, so.dptr(b), so.dptr(y))
    ref = (np.eye(n) - E) @ (Ainv @ b)
    assert np.linalg.norm(y - ref) < 1e-9 * np.linalg.norm(ref)


@pytest.mark.parametrize("pre", ["a11", "selfp", "lsc", "lsc_scaled"])
def test_schur_preconditioning_matrices_match_their_definitions(pre):
    """-pc_fieldsplit_schur_precondition a11 | selfp and PCLSC (with and without -pc_lsc_scale_diag), each with Jacobi as
    the innermost solve so that the whole preconditioner is an explicit product of known matrices."""
    p = so.Problem(6, 6, kkt=True, rhs_kind=1)
    base = ("-ksp_type fgmres -pc_type fieldsplit -pc_fieldsplit_type schur -pc_fieldsplit_schur_fact_type upper "
            "-fieldsplit_0_ksp_type preonly -fieldsplit_0_pc_type jacobi -fieldsplit_1_ksp_type preonly ")
    if pre in ("a11", "selfp"):
        opts = base + "-pc_fieldsplit_schur_precondition %s -fieldsplit_1_pc_type jacobi" % pre
    else:
        opts = base + ("-pc_fieldsplit_schur_precondition self -fieldsplit_1_pc_type lsc -fieldsplit_1_lsc_ksp_type preonly "
                       "-fieldsplit_1_lsc_pc_type jacobi" + (" -fieldsplit_1_pc_lsc_scale_diag" if pre == "lsc_scaled" else ""))
    s = so.Solver(p, opts)
    n0, n1 = p.nu, p.np_
    A, Bt, B, C = (m.scipy().toarray() for m in (p.A, p.Bt, p.B, p.C))
    Dinv = np.diag(1.0 / np.diag(A))
    if pre == "a11":
        dd = np.diag(C)
        Sinv = np.diag(1.0 / np.where(dd == 0, 1.0, dd))
    elif pre == "selfp":                                     # Sp = A11 - A10 diag(A00)^-1 A01
        dd = np.diag(C - B @ Dinv @ Bt)
        Sinv = np.diag(1.0 / np.where(dd == 0, 1.0, dd))
    else:                                                    # PCLSC: L^-1 A10 [D^-1] A00 [D^-1] A01 L^-1, L = A10 [D^-1] A01
        W = Dinv if pre == "lsc_scaled" else np.eye(n0)
        dl = np.diag(B @ W @ Bt)
        Linv = np.diag(1.0 / np.where(dl == 0, 1.0, dl))
        Sinv = Linv @ B @ W @ A @ W @ Bt @ Linv
    rng = np.random.default_rng(9)
    b = rng.standard_normal(n0 + n1)
    y1 = Sinv @ b[n0:]
    y0 = Dinv @ (b[:n0] - Bt @ y1)                           # UPPER factorisation
    y = np.empty(n0 + n1)
    so.lib().or_op_apply(s.ksp.contents.M, so.dptr(b), so.dptr(y))
    ref = np.concatenate([y0, y1])
    assert np.max(np.abs(y - ref)) < 1e-11 * np.max(np.abs(ref)), pre
