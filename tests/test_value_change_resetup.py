"""GPU test: changing matrix values after KSPSetUp.  PETSc tracks an object state on every Mat and PCSetUp runs again
inside the next KSPSolve when the operator changed; libb200sp does the same (Csr::state, Solver::current()).  Everything
derived from the values -- Jacobi diagonals, Chebyshev bounds, the per-tile value dictionaries of the SpMV kernels and
the CUDA graphs that captured their addresses -- must be rebuilt, never reused."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

OPTS = ("-ksp_type fgmres -ksp_rtol 1e-8 -pc_type fieldsplit -pc_fieldsplit_type schur -pc_fieldsplit_schur_fact_type upper "
        "-pc_fieldsplit_schur_precondition user -fieldsplit_0_ksp_type preonly -fieldsplit_0_pc_type mg -fieldsplit_0_pc_mg_levels 3 "
        "-fieldsplit_1_ksp_type preonly -fieldsplit_1_pc_type jacobi")


def test_solve_after_value_change_equals_a_fresh_solver(ctx):
    import saddle_point_petsc_b200 as sp
    nx = 32
    prob = sp.SaddlePointProblem(ctx, nx, nx, kkt=True, rhs_kind=1)
    ksp = prob.make_ksp(OPTS)
    x1 = sp.Vec(ctx, prob.n)
    r1 = ksp.solve(prob.rhs, x1)
    ksp.solve(prob.rhs, x1)                                  # second solve: preconditioner replayed from its CUDA graphs
    assert r1["reason"] == 2
    # pin a few interior velocity dofs (rows and columns zeroed, unit diagonal): the operator and its diagonal change
    M = nx + 1
    rows = np.array([2 * (M * 10 + 7), 2 * (M * 10 + 7) + 1, 2 * (M * 20 + 13), 2 * (M * 5 + 25) + 1], dtype=np.int32)
    prob.A.zero_rows_columns(rows, 1.0)
    prob.Bt.zero_rows(rows, 0.0)
    prob.B.zero_columns(rows)
    x2 = sp.Vec(ctx, prob.n)
    r2 = ksp.solve(prob.rhs, x2)                             # same KSP object: must notice the change and set up again
    fresh = prob.make_ksp(OPTS)
    x3 = sp.Vec(ctx, prob.n)
    r3 = fresh.solve(prob.rhs, x3)
    assert r2["reason"] == r3["reason"] == 2 and r2["its"] == r3["its"]
    a, b = x2.numpy(), x3.numpy()
    assert np.max(np.abs(a - b)) <= 1e-12 * np.max(np.abs(b))
    assert np.max(np.abs(a - x1.numpy())) > 1e-6 * np.max(np.abs(b))      # and the answer did change
    # the modified operator really is what was solved: true residual through an independent SpMV
    r = sp.Vec(ctx, prob.n)
    prob.K.residual(prob.rhs, x2, r)
    assert r.norm() <= 1e-7 * prob.rhs.norm()
