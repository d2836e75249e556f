"""BASELINE config 4: the 3-D Q1-hexahedron KKT problem.  The reference is 2-D only (DIM 2, include/Discretization.h:8), so the
3-D discretisation is the trilinear analogue defined in oracle/sp_oracle3d.c ("parity unpinned": there is no reference
code to pin against).  CPU tests check the definition against continuum identities; GPU tests check the CUDA assembly
bit for bit against it, SpMV, and the MINRES + block-diagonal + Chebyshev solve that config 4 names."""
import numpy as np
import pytest

import sp_oracle as so

OPTS_MINRES = ("-ksp_type minres -ksp_rtol 1e-8 -pc_type fieldsplit -pc_fieldsplit_type schur -pc_fieldsplit_schur_fact_type diag "
               "-pc_fieldsplit_schur_precondition user -fieldsplit_0_ksp_type chebyshev -fieldsplit_0_ksp_max_it 4 -fieldsplit_0_pc_type jacobi "
               "-fieldsplit_1_ksp_type preonly -fieldsplit_1_pc_type jacobi")
OPTS_FGMRES = OPTS_MINRES.replace("-ksp_type minres", "-ksp_type fgmres").replace("fact_type diag", "fact_type upper")


def same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint64), np.ascontiguousarray(b).view(np.uint64))


def fields(p):
    M, N, P = p.M, p.N, p.P
    X, Y, Z = np.meshgrid(np.linspace(0, 1, M), np.linspace(0, 1, N), np.linspace(0, 1, P), indexing="ij")

    def f(ux, uy, uz):
        u = np.zeros(3 * M * N * P)
        for c, a in enumerate((ux, uy, uz)):
            u[c::3] = (a + 0 * X).transpose(2, 1, 0).ravel()
        return u
    return X, Y, Z, f


def test_3d_element_definition_against_continuum_identities():
    p = so.Problem3D(6, 5, 4, bc=False)
    A, B, Bt, Cm, Q = (m.scipy() for m in (p.A, p.B, p.Bt, p.C, p.Q))
    X, Y, Z, f = fields(p)
    assert abs(A - A.T).max() < 1e-15
    for rigid in (f(1, 0, 0), f(0, 1, 0), f(0, 0, 1), f(-Y, X, 0), f(0, -Z, Y), f(Z, 0, -X)):
        assert abs(A @ rigid).max() < 1e-14                               # rigid-body modes carry no strain energy
    u = f(X, 0, 0)
    assert abs(u @ A @ u - 2.0) < 1e-10                                   # int 2 exx^2 = 2
    u = f(Y, 0, 0)
    assert abs(u @ A @ u - 1.0) < 1e-10                                   # int gxy^2 = 1
    assert abs(B - Bt.T).max() == 0.0
    assert abs((B @ f(X, Y, Z)).sum() + 3.0) < 1e-12                      # -int div u = -3
    assert abs(Q.sum() + 1.0) < 1e-12 and abs(Cm @ np.ones(Cm.shape[0])).max() < 1e-15
    n = (p.M, p.N, p.P)
    assert A.nnz == 9 * np.prod([3 * m - 2 for m in n])                   # DMCreateMatrix 27-point pattern, 3 x 3 dof


@pytest.mark.parametrize("opts", [OPTS_MINRES, OPTS_FGMRES])
def test_3d_kkt_solvers_reach_the_direct_solution(opts):
    import scipy.sparse.linalg as spla
    p = so.Problem3D(6, 6, 6)
    r = so.Solver(p, opts).solve()
    assert r["reason"] == 2
    K = p.scipy_K()
    assert np.linalg.norm(p.rhs - K @ r["x"]) / np.linalg.norm(p.rhs) < 5e-7
    Kp = K.tolil()                                                           # pin one pressure: the constant mode
    pin = p.nu
    Kp[pin, :] = 0.0; Kp[:, pin] = 0.0; Kp[pin, pin] = 1.0
    b = p.rhs.copy(); b[pin] = 0.0
    xd = spla.spsolve(Kp.tocsc(), b)
    assert np.max(np.abs(r["x"][:p.nu] - xd[:p.nu])) < 1e-5 * np.max(np.abs(xd[:p.nu]))


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("n", [(2, 2, 2), (5, 4, 3), (12, 9, 7)])
def test_3d_device_assembly_bit_exact(ctx, n):
    import saddle_point_petsc_b200 as sp
    dev = sp.SaddlePointProblem3D(ctx, *n)
    orc = so.Problem3D(*n)
    for name in ("A", "Bt", "B", "C", "Q"):
        rp, col, val = getattr(dev, name).csr()
        o = getattr(orc, name)
        assert np.array_equal(rp, o.rowptr) and np.array_equal(col, o.col), name
        assert same_bits(val, o.val), name
    assert same_bits(dev.rhs.numpy(), orc.rhs)
    assert np.array_equal(dev.bc, orc.bc)
    rng = np.random.default_rng(3)
    for name in ("A", "Bt", "B", "C"):                                    # MatMult of every block (row sums in CSR order or tree order)
        D, O = getattr(dev, name), getattr(orc, name)
        x = rng.uniform(-1, 1, O.ncols)
        y = sp.Vec(ctx, O.nrows)
        D.mult(sp.Vec.from_numpy(ctx, x), y)
        yo = O.mult(x)
        assert np.max(np.abs(y.numpy() - yo)) <= 1e-14 * max(1.0, np.max(np.abs(yo))), name


@pytest.mark.gpu
@pytest.mark.parametrize("name,opts", [("minres", OPTS_MINRES), ("fgmres", OPTS_FGMRES)])
def test_3d_kkt_solve_parity(ctx, name, opts):
    """MINRES + block-diagonal preconditioner with a Chebyshev/Jacobi A00 solve (config 4 as named) and the FGMRES/upper
    variant: iterations +-1, final relative residual within 1e-10, solution within rel 1e-8 of the oracle's."""
    import saddle_point_petsc_b200 as sp
    n = (16, 14, 12)
    dev = sp.SaddlePointProblem3D(ctx, *n)
    orc = so.Problem3D(*n)
    ksp = dev.make_ksp(opts)
    x = sp.Vec(ctx, dev.n)
    rd = ksp.solve(dev.rhs, x)
    ro = so.Solver(orc, opts).solve()
    assert rd["reason"] == ro["reason"] == 2, (rd["reason"], ro["reason"])
    assert abs(rd["its"] - ro["its"]) <= 1, (rd["its"], ro["its"])
    xs, nu = x.numpy(), dev.nu
    if rd["its"] == ro["its"]:
        assert abs(rd["rnorm"] / rd["history"][0] - ro["rnorm"] / ro["history"][0]) <= 1e-10
        assert np.max(np.abs(xs[:nu] - ro["x"][:nu])) <= 1e-8 * np.max(np.abs(ro["x"][:nu]))
        dp = xs[nu:] - ro["x"][nu:]
        assert np.max(np.abs(dp - dp.mean())) <= 1e-8 * np.max(np.abs(ro["x"][nu:]))
    K = orc.scipy_K()
    assert np.linalg.norm(orc.rhs - K @ xs) / np.linalg.norm(orc.rhs) < 5e-7


@pytest.mark.gpu
@pytest.mark.parametrize("size,n", [(2, (8, 7, 9)), (4, (10, 9, 8)), (8, (11, 10, 9))])
def test_3d_row_partitioned_assembly_spmv_and_solve(size, n):
    """3-D DMDA partition (PETSC_DECIDE process grid, rank-contiguous numbering, ghost layer through the general halo): every
    rank's rows equal the oracle's rows (values bit for bit), distributed MatMult, and the MINRES solve of config 4."""
    import scipy.sparse as sps
    import saddle_point_petsc_b200 as sp
    orc = so.Problem3D(*n)
    M, N, P = orc.M, orc.N, orc.P
    ro = so.Solver(orc, OPTS_MINRES).solve()
    rng = np.random.default_rng(7)
    xg = {1: rng.uniform(-1, 1, M * N * P), 3: rng.uniform(-1, 1, 3 * M * N * P)}
    blocks = {"A": (3, 3), "Bt": (3, 1), "B": (1, 3), "C": (1, 1)}

    def rank_fn(ctx):
        prob = sp.SaddlePointProblem3D(ctx, *n)
        da = prob.da
        # natural ids of my owned nodes, in local order (x fastest, then y, then z)
        kk, jj, ii = np.meshgrid(np.arange(da.zs, da.zs + da.zm), np.arange(da.ys, da.ys + da.ym), np.arange(da.xs, da.xs + da.xm), indexing="ij")
        nat = ((kk * N + jj) * M + ii).ravel()
        out = {"nat": nat, "g0": da.gstart}
        for name, (dr, dc) in blocks.items():
            m = getattr(prob, name)
            x = sp.Vec.from_numpy(ctx, xg[dc].reshape(-1, dc)[nat].ravel())
            y = sp.Vec(ctx, len(nat) * dr)
            m.mult(x, y)
            out["y" + name] = y.numpy()
            out["csr" + name] = m.csr()          # local rows, GLOBAL (PETSc numbering) columns
        ksp = prob.make_ksp(OPTS_MINRES)
        x = sp.Vec(ctx, prob.n)
        r = ksp.solve(prob.rhs, x)
        out["its"], out["reason"], out["x"] = r["its"], r["reason"], x.numpy()
        out["rhs"] = prob.rhs.numpy()
        return out

    res = sp.run_ranks(size, rank_fn)
    petsc_of_nat = np.zeros(M * N * P, dtype=np.int64)                     # natural node -> PETSc global node
    for o in res:
        petsc_of_nat[o["nat"]] = o["g0"] + np.arange(len(o["nat"]))
    nat_of_petsc = np.argsort(petsc_of_nat)
    for o in res:
        nat = o["nat"]
        for name, (dr, dc) in blocks.items():
            ref = getattr(orc, name).scipy()
            rows = (nat[:, None] * dr + np.arange(dr)).ravel()
            yr = (ref @ xg[dc])[rows]
            assert np.max(np.abs(o["y" + name] - yr)) <= 1e-14 * max(1.0, np.max(np.abs(yr))), name
            rp, col, val = o["csr" + name]
            mine = sps.csr_matrix((val, col, rp), shape=(len(rows), dc * M * N * P)).tocoo()
            natcol = nat_of_petsc[mine.col // dc] * dc + mine.col % dc     # back to natural columns
            got = sps.csr_matrix((mine.data, (mine.row, natcol)), shape=mine.shape)
            got.sort_indices()
            R = ref[rows]
            assert np.array_equal(got.indptr, R.indptr) and np.array_equal(got.indices, R.indices), name
            assert same_bits(got.data, R.data), name
        assert same_bits(o["rhs"][:3 * len(nat)], orc.f.reshape(-1, 3)[nat].ravel())
    assert len({o["its"] for o in res}) == 1 and all(o["reason"] == 2 for o in res)
    assert abs(res[0]["its"] - ro["its"]) <= 1, (res[0]["its"], ro["its"])
    if res[0]["its"] == ro["its"]:
        umax = np.max(np.abs(ro["x"][:orc.nu]))
        for o in res:
            nat = o["nat"]
            xu = ro["x"][:orc.nu].reshape(-1, 3)[nat].ravel()
            assert np.max(np.abs(o["x"][:3 * len(nat)] - xu)) <= 1e-8 * umax
