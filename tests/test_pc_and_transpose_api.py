"""GPU tests of two C-ABI entry points a PETSc plugin binds besides KSP: the PC as an object of its own
(PCSetOperators / PCSetFromOptions / PCSetUp / PCApply) and MatMultTranspose."""
import numpy as np
import pytest

import sp_oracle as so
from test_oracle import CONFIGS

pytestmark = pytest.mark.gpu


def same_bits(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))


def test_mat_mult_transpose_adds_in_row_order(ctx):
    import saddle_point_petsc_b200 as sp
    nx, ny = 40, 31
    dev = sp.SaddlePointProblem(ctx, nx, ny, kkt=True)
    orc = so.Problem(nx, ny, kkt=True)
    rng = np.random.default_rng(2)
    for name in ("Bt", "B", "A"):
        D, O = getattr(dev, name), getattr(orc, name)
        x = rng.uniform(-1, 1, O.nrows)
        xd, yd = sp.Vec.from_numpy(ctx, x), sp.Vec(ctx, O.ncols)
        D.mult_transpose(xd, yd)
        assert same_bits(yd.numpy(), O.transpose().mult(x)), name      # oracle: explicit transpose, CSR-order sums
    # B was assembled as the exact transpose of B^T, so B^T-transpose-times-x is B x to the last bit
    x = rng.uniform(-1, 1, orc.Bt.nrows)
    xd, y1, y2 = sp.Vec.from_numpy(ctx, x), sp.Vec(ctx, orc.Bt.ncols), sp.Vec(ctx, orc.Bt.ncols)
    dev.Bt.mult_transpose(xd, y1)
    dev.B.mult(xd, y2)
    assert same_bits(y1.numpy(), y2.numpy())
    # the cached transpose follows value changes
    rows = np.array([0, 3, 2 * (41 * 7 + 5)], dtype=np.int32)      # two boundary dofs and an interior one
    dev.Bt.zero_rows(rows, 0.0)
    dev.Bt.mult_transpose(xd, y1)
    x0 = x.copy()
    x0[rows] = 0.0                                            # zeroed rows of B^T contribute nothing to (B^T)^T x
    assert same_bits(y1.numpy(), orc.Bt.transpose().mult(x0))
    with pytest.raises(sp.B200spError):
        dev.Bt.mult_transpose(y1, xd)                         # wrong sizes


@pytest.mark.parametrize("name", ["fgmres_upper_mg", "gmres_full_jacobi", "minres_diag_cheb"])
def test_pc_object_equals_the_pc_inside_a_ksp(ctx, name):
    import saddle_point_petsc_b200 as sp
    nx = 16
    dev = sp.SaddlePointProblem(ctx, nx, nx, kkt=True, rhs_kind=1)
    orc = so.Problem(nx, nx, kkt=True, rhs_kind=1)
    pc = sp.PC(ctx)
    pc.set_operators(dev.K, dev.K)
    pc.set_schur_user_mat(dev.Q)
    pc.set_dmda(dev.da)
    pc.set_options(CONFIGS[name])
    pc.setup()
    assert "fieldsplit" in pc.view()
    ksp = dev.make_ksp(CONFIGS[name])
    ksp.setup()
    v = np.random.default_rng(5).uniform(-1, 1, dev.n)
    vd, y1, y2 = sp.Vec.from_numpy(ctx, v), sp.Vec(ctx, dev.n), sp.Vec(ctx, dev.n)
    pc.apply(vd, y1)
    ksp.pc_apply(vd, y2)
    assert same_bits(y1.numpy(), y2.numpy())                  # same setup, same kernels
    s = so.Solver(orc, CONFIGS[name])
    yo = np.empty(dev.n)
    so.lib().or_op_apply(s.ksp.contents.M, so.dptr(v), so.dptr(yo))
    assert np.max(np.abs(y1.numpy() - yo)) <= 1e-10 * np.max(np.abs(yo))
    pc.destroy()


def test_pc_none_is_the_identity(ctx):
    import saddle_point_petsc_b200 as sp
    dev = sp.SaddlePointProblem(ctx, 8, 8)
    pc = sp.PC(ctx)
    pc.set_operators(dev.K, dev.K)
    pc.set_options("-pc_type none")
    v = np.arange(dev.n, dtype=np.float64)
    vd, yd = sp.Vec.from_numpy(ctx, v), sp.Vec(ctx, dev.n)
    pc.apply(vd, yd)                                          # sets up on first use
    assert same_bits(yd.numpy(), v)
    pc.destroy()


def test_ksp_shares_ownership_of_its_operators(ctx):
    """PETSc reference-counts the operators of KSPSetOperators: MatDestroy before KSPSolve is legal."""
    import numpy as np
    import saddle_point_petsc_b200 as sp
    opts = ("-ksp_type fgmres -ksp_rtol 1e-8 -pc_type fieldsplit -pc_fieldsplit_type schur -pc_fieldsplit_schur_fact_type upper "
            "-pc_fieldsplit_schur_precondition user -fieldsplit_0_ksp_type preonly -fieldsplit_0_pc_type mg -fieldsplit_0_pc_mg_levels 2 "
            "-fieldsplit_1_ksp_type preonly -fieldsplit_1_pc_type jacobi")
    dev = sp.SaddlePointProblem(ctx, 16, 16, kkt=True, rhs_kind=1)
    ksp = dev.make_ksp(opts)
    x = sp.Vec(ctx, dev.n)
    r1 = ksp.solve(dev.rhs, x)
    x1 = x.numpy()
    ksp2 = dev.make_ksp(opts)
    for m in (dev.K, dev.A, dev.Bt, dev.B, dev.C, dev.Q):     # every caller-side handle goes away before the second KSP is even set up
        m.destroy()
    r2 = ksp2.solve(dev.rhs, x)
    assert r2["its"] == r1["its"] and np.array_equal(x.numpy(), x1)


def test_options_that_would_silently_change_the_solver_are_errors(ctx):
    """-options_left promoted to an error, PETSc defaults this library does not have, unsupported sides / norms."""
    import pytest
    import saddle_point_petsc_b200 as sp
    dev = sp.SaddlePointProblem(ctx, 8, 8)
    for bad in ("-ksp_type gmres",                                        # no -pc_type: PETSc would use ILU(0)
                "-ksp_type gmres -pc_type jacobi -ksp_gmres_restrat 5",   # mistyped option
                "-ksp_type gmres -pc_type jacobi -ksp_pc_side right",     # GMRES is implemented left-preconditioned only
                "-ksp_type fgmres -pc_type jacobi -ksp_norm_type preconditioned",
                "-ksp_type gmres -pc_type jacobi -pc_mg_levels 3",        # option of a PC that is not in use
                "-ksp_type gmres -pc_type mg -pc_mg_levels 2 -mg_levels_pc_type sor"):
        with pytest.raises(sp.B200spError) as e:
            dev.make_ksp(bad).setup()
        assert e.value.code == 4, (bad, str(e.value))                     # B200SP_ERR_UNSUPPORTED
    dev.make_ksp("-ksp_type gmres -pc_type jacobi -ksp_monitor -ksp_converged_reason -ksp_pc_side left -ksp_norm_type preconditioned").setup()
