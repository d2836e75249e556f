"""CPU tests of the oracle against the known answers derived for the reference's discretisation
 and against scipy.  No GPU."""
import numpy as np
import pytest

import sp_oracle as so

FREE = [10, 11, 12, 13, 18, 19, 20, 21]
U_FREE = [0.0496585971446208, 0.0918684047175528, 0.0397268777157024, 0.0869025450030936,
          0.0397268777157024, 0.0869025450030936, 0.0496585971446208, 0.0918684047175528]
KE = np.array([[1, .25, 0, .25, -.5, -.25, -.5, -.25], [.25, 1, -.25, -.5, -.25, -.5, .25, 0],
               [0, -.25, 1, -.25, -.5, .25, -.5, .25], [.25, -.5, -.25, 1, -.25, 0, .25, -.5],
               [-.5, -.25, -.5, -.25, 1, .25, 0, .25], [-.25, -.5, .25, 0, .25, 1, -.25, -.5],
               [-.5, .25, -.5, .25, 0, -.25, 1, -.25], [-.25, 0, .25, -.5, .25, -.5, -.25, 1]])


def test_element_matrix_known_answer():
    ke = so.element_stress(so.element_coords(4, 4, 1, 1)).reshape(8, 8)
    # the truncated Gauss literal 0.57735026919 moves the entries by ~1e-11 from the rationals
    assert np.abs(ke - KE).max() < 1e-10
    assert np.abs(ke - KE).max() > 0.0
    assert np.abs(ke - ke.T).max() < 1e-15
    assert np.abs(ke.sum(axis=1)).max() < 1e-14


def test_element_rhs_known_answer():
    h = 1.0 / 3.0
    fe = so.element_rhs(so.element_coords(4, 4, 0, 0)).reshape(4, 2)
    assert np.allclose(fe, (h * h / 4) * np.array([1.0, 2.0]), rtol=1e-10)


def test_default_grid_structure_and_solution():
    p = so.Problem(3, 3)
    assert p.A.nnz == 400 and p.A.nrows == 32
    assert np.bincount(np.diff(p.A.rowptr))[[8, 12, 18]].tolist() == [8, 16, 8]
    assert len(p.bc) == 24
    assert sorted(set(range(32)) - set(p.bc.tolist())) == FREE
    r10 = p.A.col[p.A.rowptr[10]:p.A.rowptr[11]].tolist()
    assert r10 == [0, 1, 2, 3, 4, 5, 8, 9, 10, 11, 12, 13, 16, 17, 18, 19, 20, 21]
    K = p.A.scipy().toarray()
    u = np.linalg.solve(K, p.f)
    assert np.allclose(u[FREE], U_FREE, rtol=1e-9)
    assert np.abs(u[p.bc]).max() == 0.0
    assert np.allclose(np.diag(K)[FREE], 4.0, rtol=1e-10)


@pytest.mark.parametrize("m", [4, 17, 65])
def test_closed_form_counts(m):
    p = so.Problem(m - 1, m - 1)
    assert p.A.nnz == 4 * (3 * m - 2) ** 2
    assert len(p.bc) == 2 * (4 * m - 4)


def test_as_written_mode_documents_the_defect():
    p = so.Problem(3, 3, as_written=True)
    nan_rows = sorted({r for r in range(32) for k in range(p.A.rowptr[r], p.A.rowptr[r + 1]) if np.isnan(p.A.val[k])})
    assert nan_rows == FREE and int(np.isnan(p.A.val).sum()) == 64
    assert np.abs(p.f).max() == 0.0


def test_dmda_partition_rules():
    assert so.dmda_proc_grid(2310, 2310, 1) == (1, 1)
    assert so.dmda_proc_grid(2310, 2310, 2) == (1, 2)
    assert so.dmda_proc_grid(2310, 2310, 4) == (2, 2)
    assert so.dmda_proc_grid(2310, 2310, 8) == (2, 4)
    assert so.dmda_ownership(2310, 4).tolist() == [578, 578, 577, 577]
    nm, ow = so.dmda_natural_to_petsc(7, 5, 4)
    assert sorted(nm.tolist()) == list(range(35))
    # rank-contiguous numbering
    for r in range(4):
        ids = np.sort(nm[ow == r])
        assert ids.tolist() == list(range(ids[0], ids[0] + len(ids)))
    assert so.dmda_element_range(7, 5, 1, 0) == (0, 0, 6, 4)


def test_kkt_blocks_properties():
    p = so.Problem(6, 5, kkt=True, rhs_kind=1)
    Bt, B = p.Bt.scipy(), p.B.scipy()
    assert abs(B - Bt.T).max() == 0.0                       # divergence is the exact transpose of the gradient
    ones = np.ones(p.np_)
    assert np.abs(p.C.scipy() @ ones).max() < 1e-16         # stabilisation annihilates constants
    assert np.abs(Bt @ ones).max() < 1e-15                  # constant pressure null vector (enclosed flow)
    Q = p.Q.scipy()
    assert np.isclose(-(ones @ (Q @ ones)), 1.0, rtol=1e-12)  # -Q is the mass matrix of the unit square
    K = p.scipy_K().toarray()
    assert np.abs(K - K.T).max() < 1e-15


def test_interp_matches_transpose_and_partition_of_unity():
    P = so.Csr(so.lib().or_interp_q1(5, 4, 2, 0)).scipy()
    assert P.shape == (2 * 9 * 7, 2 * 5 * 4)
    assert np.allclose(P @ np.ones(P.shape[1]), 1.0)
    Pb = so.Csr(so.lib().or_interp_q1(5, 4, 2, 1)).scipy()
    assert (Pb != 0).sum() < (P != 0).sum()


def test_matmat_and_transpose_against_scipy():
    p = so.Problem(5, 4, kkt=True)
    L = p.B.matmat(p.Bt).scipy()
    ref = (p.B.scipy() @ p.Bt.scipy()).tocsr()
    assert abs(L - ref).max() < 1e-16
    assert abs(p.Bt.transpose().scipy() - p.Bt.scipy().T).max() == 0.0


OPTS_BASE = ("-ksp_rtol 1e-8 -pc_type fieldsplit -pc_fieldsplit_type schur -pc_fieldsplit_schur_precondition user "
             "-fieldsplit_1_ksp_type preonly -fieldsplit_1_pc_type jacobi ")
CONFIGS = {
    "gmres_full_jacobi": "-ksp_type gmres -pc_fieldsplit_schur_fact_type full -fieldsplit_0_ksp_type preonly -fieldsplit_0_pc_type jacobi " + OPTS_BASE,
    "fgmres_upper_mg": "-ksp_type fgmres -pc_fieldsplit_schur_fact_type upper -fieldsplit_0_ksp_type preonly -fieldsplit_0_pc_type mg -fieldsplit_0_pc_mg_levels 3 " + OPTS_BASE,
    "minres_diag_cheb": "-ksp_type minres -pc_fieldsplit_schur_fact_type diag -fieldsplit_0_ksp_type chebyshev -fieldsplit_0_ksp_max_it 4 -fieldsplit_0_pc_type jacobi " + OPTS_BASE,
    "gmres_lower_selfp": "-ksp_type gmres -ksp_rtol 1e-8 -pc_type fieldsplit -pc_fieldsplit_schur_fact_type lower -pc_fieldsplit_schur_precondition selfp -fieldsplit_0_ksp_type chebyshev -fieldsplit_0_ksp_max_it 3 -fieldsplit_0_pc_type jacobi -fieldsplit_1_ksp_type preonly -fieldsplit_1_pc_type jacobi",
    "fgmres_lsc": "-ksp_type fgmres -ksp_rtol 1e-8 -pc_type fieldsplit -pc_fieldsplit_schur_fact_type upper -pc_fieldsplit_schur_precondition self -fieldsplit_0_ksp_type preonly -fieldsplit_0_pc_type mg -fieldsplit_0_pc_mg_levels 2 -fieldsplit_1_ksp_type preonly -fieldsplit_1_pc_type lsc -fieldsplit_1_pc_lsc_scale_diag -fieldsplit_1_lsc_ksp_type chebyshev -fieldsplit_1_lsc_ksp_max_it 8 -fieldsplit_1_lsc_pc_type jacobi",
}


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_kkt_solvers_converge_to_the_direct_solution(name):
    import scipy.sparse.linalg as spla
    p = so.Problem(16, 16, kkt=True, rhs_kind=1)
    r = so.Solver(p, CONFIGS[name]).solve()
    assert r["reason"] == 2, (name, r["reason"], r["its"])
    K = p.scipy_K()
    res = np.linalg.norm(p.rhs - K @ r["x"]) / np.linalg.norm(p.rhs)
    assert res < 5e-7, (name, res)
    # velocity is unique; pressure is unique up to a constant (enclosed flow)
    Kp = K.tolil()
    Kp[p.nu, :] = 0.0
    Kp[p.nu, p.nu] = 1.0
    rhs = p.rhs.copy()
    rhs[p.nu] = 0.0
    xd = spla.spsolve(Kp.tocsc(), rhs)
    assert np.abs(r["x"][:p.nu] - xd[:p.nu]).max() < 1e-6 * max(1.0, np.abs(xd[:p.nu]).max())


def test_velocity_only_problem_all_krylov_types():
    p = so.Problem(12, 9)
    u = np.linalg.solve(p.A.scipy().toarray(), p.f)
    for opts in ("-ksp_type gmres -pc_type jacobi", "-ksp_type fgmres -pc_type jacobi", "-ksp_type minres -pc_type jacobi",
                 "-ksp_type gmres -pc_type none", "-ksp_type fgmres -pc_type mg -pc_mg_levels 2"):
        if "mg" in opts:
            p2 = so.Problem(12, 8)
            u2 = np.linalg.solve(p2.A.scipy().toarray(), p2.f)
            r = so.Solver(p2, opts + " -ksp_rtol 1e-10").solve()
            assert r["reason"] == 2 and np.abs(r["x"] - u2).max() < 1e-8
            assert r["its"] < 15
        else:
            r = so.Solver(p, opts + " -ksp_rtol 1e-10").solve()
            assert r["reason"] == 2 and np.abs(r["x"] - u).max() < 1e-8, opts


# ------------------------------------------------------------------ the reference's own constraint block (4 dense rows)
G_CON = (0.01, -0.02, 0.005, 0.003)
CON_OPTS = ("-ksp_rtol 1e-10 -pc_type fieldsplit -pc_fieldsplit_type schur -pc_fieldsplit_schur_precondition selfp "
            "-fieldsplit_0_ksp_type preonly -fieldsplit_0_pc_type mg -fieldsplit_0_pc_mg_levels 3 -fieldsplit_1_ksp_type preonly -fieldsplit_1_pc_type lu ")
CON_CONFIGS = {"fgmres_upper": "-ksp_type fgmres -pc_fieldsplit_schur_fact_type upper " + CON_OPTS,
               "minres_diag": "-ksp_type minres -pc_fieldsplit_schur_fact_type diag " + CON_OPTS,
               "gmres_full": "-ksp_type gmres -pc_fieldsplit_schur_fact_type full " + CON_OPTS}


def test_constraint_rows_are_the_moments_they_claim_to_be():
    """B is 4 x nCols (src/SaddlePointProblem.c:48-49): barycentre x / y, dilation and rotation moments about the centre."""
    nx, ny = 12, 9
    p = so.Problem(nx, ny, constraints=True, bc=False)
    B = p.B.scipy()
    M, N = nx + 1, ny + 1
    assert B.shape == (4, 2 * M * N) and abs(p.Bt.scipy() - B.T).max() == 0.0
    X, Y = np.meshgrid(np.linspace(0, 1, M), np.linspace(0, 1, N))
    def field(ux, uy):
        u = np.zeros(2 * M * N)
        u[0::2], u[1::2] = ux.ravel(), uy.ravel()
        return u
    one, zero = np.ones_like(X), np.zeros_like(X)
    assert np.allclose(B @ field(one, zero), [1.0, 0.0, 0.0, 0.0], atol=1e-12)          # int 1 = area
    assert np.allclose(B @ field(zero, one), [0.0, 1.0, 0.0, 0.0], atol=1e-12)
    assert np.allclose(B @ field(X - 0.5, Y - 0.5), [0.0, 0.0, 1.0 / 6.0, 0.0], atol=1e-9)   # int |x-c|^2 = 2/12 (Gauss literal: 1e-11)
    assert np.allclose(B @ field(-(Y - 0.5), X - 0.5), [0.0, 0.0, 0.0, 1.0 / 6.0], atol=1e-9)  # rigid rotation


@pytest.mark.parametrize("name", sorted(CON_CONFIGS))
def test_constrained_problem_solves_to_the_direct_solution(name):
    """[A Bt; B 0][u; lambda] = [f; g] (the commented-out wiring of src/SaddlePointProblem.c:45-60): Schur complement is the
    dense 4 x 4 -B A^-1 Bt, preconditioned by selfp (-B diag(A)^-1 Bt) solved exactly."""
    import scipy.sparse.linalg as spla
    p = so.Problem(16, 16, constraints=True, g=G_CON)
    x = spla.spsolve(p.scipy_K().tocsc(), p.rhs)
    assert np.allclose(p.B.scipy() @ x[:-4], G_CON, atol=1e-13)
    r = so.Solver(p, CON_CONFIGS[name]).solve()
    assert r["reason"] == 2 and r["its"] <= 25, (r["reason"], r["its"])
    assert np.max(np.abs(r["x"] - x)) <= 1e-8 * np.max(np.abs(x))
