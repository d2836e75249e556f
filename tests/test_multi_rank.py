"""Row-partitioned (DMDA) path: every rank is a thread of this process driving its own context on cuda:0
(b200sp LocalGroup: collectives are host barriers + device copies, the ranks never wait on each other inside a
kernel), so the whole distributed algorithm -- rank-local assembly with ghost elements, MPIAIJ-style diag/off-diag
split, halo exchange, global reductions, distributed multigrid -- is checked on a 1-GPU box against the oracle.
The NCCL transport used by bench.py under torchrun shares everything except the Comm object."""
import numpy as np
import pytest
import scipy.sparse as sps

import sp_oracle as so

pytestmark = pytest.mark.gpu


def petsc_perm(M, N, size, dof):
    nm, ow = so.dmda_natural_to_petsc(M, N, size)
    return (np.repeat(nm.astype(np.int64) * dof, dof) + np.tile(np.arange(dof), M * N)), nm, ow


def to_petsc_order(A, M, N, size, dofr, dofc):
    pr, _, _ = petsc_perm(M, N, size, dofr)
    pc, _, _ = petsc_perm(M, N, size, dofc)
    C = A.scipy().tocoo()
    P = sps.csr_matrix((C.data, (pr[C.row], pc[C.col])), shape=C.shape)   # explicit zeros are kept by csr_matrix((data,(i,j)))
    P.sort_indices()
    return P


def same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint64), np.ascontiguousarray(b).view(np.uint64))


@pytest.mark.parametrize("size,nx,ny", [(2, 9, 7), (4, 12, 10), (8, 21, 17), (3, 14, 5)])
def test_distributed_assembly_and_spmv_match_the_oracle(size, nx, ny):
    import saddle_point_petsc_b200 as sp
    M, N = nx + 1, ny + 1
    orc = so.Problem(nx, ny, kkt=True, rhs_kind=1)
    blocks = {"A": (2, 2), "Bt": (2, 1), "B": (1, 2), "C": (1, 1), "Q": (1, 1)}
    ref = {k: to_petsc_order(getattr(orc, k), M, N, size, *d) for k, d in blocks.items()}
    pu, nm, ow = petsc_perm(M, N, size, 2)
    pp, _, _ = petsc_perm(M, N, size, 1)
    rng = np.random.default_rng(5)
    xg = {1: rng.uniform(-1, 1, M * N), 2: rng.uniform(-1, 1, 2 * M * N)}      # global vectors in PETSc order
    f_petsc = np.zeros(2 * M * N)
    f_petsc[pu] = orc.f

    def rank_fn(ctx):
        r = ctx.rank
        prob = sp.SaddlePointProblem(ctx, nx, ny, kkt=True, rhs_kind=1)
        nl = prob.da.n_nodes_local
        g0 = int(np.sum(ow < r))                     # rank-contiguous numbering: first global node of this rank
        assert (prob.da.xs, prob.da.ys, prob.da.xm, prob.da.ym) == sp.dmda_corners(M, N, size, r)
        out = {}
        for name, (dr, dc) in blocks.items():
            m = getattr(prob, name)
            rp, col, val = m.csr()                   # local rows, GLOBAL PETSc column ids
            R = ref[name][g0 * dr:(g0 + nl) * dr]
            assert np.array_equal(rp, R.indptr), (name, r)
            assert np.array_equal(col, R.indices), (name, r)
            assert same_bits(val, R.data), (name, r)
            # distributed MatMult: halo exchange + diagonal block + off-diagonal block
            x = sp.Vec.from_numpy(ctx, xg[dc][g0 * dc:(g0 + nl) * dc])
            y = sp.Vec(ctx, nl * dr)
            m.mult(x, y)
            yr = ref[name] @ xg[dc]
            out[name] = np.max(np.abs(y.numpy() - yr[g0 * dr:(g0 + nl) * dr])) / max(1.0, np.max(np.abs(yr)))
        assert same_bits(prob.rhs.numpy()[:2 * nl], f_petsc[g0 * 2:(g0 + nl) * 2])
        # global reductions
        v = sp.Vec.from_numpy(ctx, xg[2][g0 * 2:(g0 + nl) * 2])
        out["norm"] = v.norm()
        return out

    res = sp.run_ranks(size, rank_fn)
    for o in res:
        for name in blocks:
            assert o[name] < 1e-14, (name, o[name])
        assert abs(o["norm"] - np.linalg.norm(xg[2])) < 1e-12 * np.linalg.norm(xg[2])
    assert len({o["norm"] for o in res}) == 1        # every rank holds the same reduced value


CFG_CHEB = ("-ksp_type fgmres -ksp_rtol 1e-8 -pc_type fieldsplit -pc_fieldsplit_type schur -pc_fieldsplit_schur_fact_type upper "
            "-pc_fieldsplit_schur_precondition user -fieldsplit_0_ksp_type chebyshev -fieldsplit_0_ksp_max_it 6 -fieldsplit_0_pc_type jacobi "
            "-fieldsplit_1_ksp_type preonly -fieldsplit_1_pc_type jacobi")
CFG_MINRES = ("-ksp_type minres -ksp_rtol 1e-8 -pc_type fieldsplit -pc_fieldsplit_type schur -pc_fieldsplit_schur_fact_type diag "
              "-pc_fieldsplit_schur_precondition user -fieldsplit_0_ksp_type chebyshev -fieldsplit_0_ksp_max_it 4 -fieldsplit_0_pc_type jacobi "
              "-fieldsplit_1_ksp_type preonly -fieldsplit_1_pc_type jacobi")
CFG_MG = ("-ksp_type fgmres -ksp_rtol 1e-8 -pc_type fieldsplit -pc_fieldsplit_type schur -pc_fieldsplit_schur_fact_type upper "
          "-pc_fieldsplit_schur_precondition user -fieldsplit_0_ksp_type preonly -fieldsplit_0_pc_type mg -fieldsplit_0_pc_mg_levels 4 "
          "-fieldsplit_1_ksp_type preonly -fieldsplit_1_pc_type jacobi")


@pytest.mark.parametrize("size", [2, 4, 8])
@pytest.mark.parametrize("cfg", ["cheb", "minres", "mg"])
def test_distributed_kkt_solve_matches_single_rank_and_oracle(size, cfg):
    import saddle_point_petsc_b200 as sp
    nx = ny = 32 if cfg != "mg" else 48
    opts = {"cheb": CFG_CHEB, "minres": CFG_MINRES, "mg": CFG_MG}[cfg]
    M, N = nx + 1, ny + 1
    orc = so.Problem(nx, ny, kkt=True, rhs_kind=1)
    ro = so.Solver(orc, opts).solve()
    pu, nm, ow = petsc_perm(M, N, size, 2)
    pp, _, _ = petsc_perm(M, N, size, 1)
    xu = np.zeros(2 * M * N); xu[pu] = ro["x"][:2 * M * N]
    xp = np.zeros(M * N); xp[pp] = ro["x"][2 * M * N:]

    def rank_fn(ctx):
        prob = sp.SaddlePointProblem(ctx, nx, ny, kkt=True, rhs_kind=1)
        ksp = prob.make_ksp(opts)
        x = sp.Vec(ctx, prob.n)
        r = ksp.solve(prob.rhs, x)
        nl = prob.da.n_nodes_local
        g0 = int(np.sum(ow < ctx.rank))
        xs = x.numpy()
        eu = np.max(np.abs(xs[:2 * nl] - xu[2 * g0:2 * (g0 + nl)]))
        dp = xs[2 * nl:] - xp[g0:g0 + nl]
        rr = sp.Vec(ctx, prob.n)
        prob.K.residual(prob.rhs, x, rr)
        return r["its"], r["reason"], eu, dp, rr.norm() / prob.rhs.norm()

    res = sp.run_ranks(size, rank_fn)
    its = {r[0] for r in res}
    assert len(its) == 1 and all(r[1] == 2 for r in res)
    assert abs(res[0][0] - ro["its"]) <= 1, (res[0][0], ro["its"])
    # the distributed solve's own true residual meets the tolerance whatever the iteration count ...
    assert res[0][4] < 5e-7, res[0][4]
    if res[0][0] == ro["its"]:
        # ... and with the same number of iterations the iterates agree with the single-rank oracle to north_star's
        # 1e-8 (the partition only changes summation orders: measured differences are 1e-10 ... 1e-13).  The weakly
        # preconditioned configuration (Chebyshev(6)/Jacobi instead of multigrid: ~10x the iterations) is more
        # sensitive to rounding than that: its tolerance is the ORACLE'S OWN measured sensitivity to a mathematically
        # neutral change (modified instead of classical Gram-Schmidt), as for the LSC configuration in test_gpu_parity.
        tol_u = tol_p = 1e-8
        if cfg == "cheb":
            # measured: 1e-8 ... 6e-8 between two valid rtol-1e-8 iterates here.  Bound: the distributed iterate may not be
            # farther from the oracle's iterate than the oracle's iterate is from the converged solution (rtol 1e-13).
            rt = so.Solver(orc, opts.replace("-ksp_rtol 1e-8", "-ksp_rtol 1e-13")).solve()
            nu_ = 2 * M * N
            d2 = rt["x"][nu_:] - ro["x"][nu_:]
            tol_u = max(tol_u, np.max(np.abs(rt["x"][:nu_] - ro["x"][:nu_])) / np.max(np.abs(xu)))
            tol_p = max(tol_p, np.max(np.abs(d2 - d2.mean())) / np.max(np.abs(xp)))
        assert max(r[2] for r in res) <= tol_u * np.max(np.abs(xu)), (max(r[2] for r in res) / np.max(np.abs(xu)), tol_u)
        dp = np.concatenate([r[3] for r in res])
        assert np.max(np.abs(dp - dp.mean())) <= tol_p * np.max(np.abs(xp)), (np.max(np.abs(dp - dp.mean())) / np.max(np.abs(xp)), tol_p)
    else:
        # one iteration more or less at rtol 1e-8: both are solutions to the tolerance, they differ by about the
        # last correction (measured: a few 1e-7 of the solution)
        assert max(r[2] for r in res) <= 5e-6 * np.max(np.abs(xu))


# ------------------------------------------------------------------ distributed SpGEMM: selfp and LSC on a row-partitioned nest
@pytest.mark.parametrize("size,nx,ny", [(2, 12, 9), (4, 14, 12), (8, 24, 20)])
def test_distributed_matmatmult_matches_the_oracle(size, nx, ny):
    """L = B * Bt and Sp-style products on row-partitioned operands: the rows each rank holds equal the oracle's product
    (structure exactly, in PETSc numbering; values to rounding -- the products of an entry are added in local column order)."""
    import saddle_point_petsc_b200 as sp
    M, N = nx + 1, ny + 1
    orc = so.Problem(nx, ny, kkt=True, rhs_kind=1)
    Lref = to_petsc_order(orc.B.matmat(orc.Bt), M, N, size, 1, 1)
    Lref.eliminate_zeros()                                            # structural zeros may differ between the two orders of construction
    pp, nm, ow = petsc_perm(M, N, size, 1)
    rng = np.random.default_rng(9)
    xg = rng.uniform(-1, 1, M * N)

    def rank_fn(ctx):
        prob = sp.SaddlePointProblem(ctx, nx, ny, kkt=True, rhs_kind=1)
        nl = prob.da.n_nodes_local
        g0 = int(np.sum(ow < ctx.rank))
        Lm = prob.B.matmult(prob.Bt)
        rp, col, val = Lm.csr()                                       # local rows, global PETSc columns
        mine = sps.csr_matrix((val, col, rp), shape=(nl, M * N))
        mine.eliminate_zeros()
        R = Lref[g0:g0 + nl]
        d = abs(mine - R)
        err = d.max() / abs(Lref).max() if d.nnz else 0.0
        same_pattern = np.array_equal(mine.indptr, R.indptr) and np.array_equal(mine.indices, R.indices)
        x = sp.Vec.from_numpy(ctx, xg[g0:g0 + nl])
        y = sp.Vec(ctx, nl)
        for rep in range(2):                                           # the product's own (two-node-wide) halo, both parities
            Lm.mult(x, y)
        yr = Lref @ xg
        return err, same_pattern, np.max(np.abs(y.numpy() - yr[g0:g0 + nl])) / np.max(np.abs(yr))

    for err, same_pattern, merr in sp.run_ranks(size, rank_fn):
        assert same_pattern
        assert err < 1e-14 and merr < 1e-13, (err, merr)


@pytest.mark.parametrize("size", [2, 4])
@pytest.mark.parametrize("name", ["gmres_lower_selfp", "fgmres_lsc"])
def test_distributed_selfp_and_lsc_match_the_oracle(size, name):
    """-pc_fieldsplit_schur_precondition selfp and -fieldsplit_1_pc_type lsc on a row-partitioned nest (the BASELINE
    'FGMRES + Schur/LSC on 1/2/4/8 GPUs' configuration): one preconditioner application and the solve against the oracle."""
    import saddle_point_petsc_b200 as sp
    from test_oracle import CONFIGS
    nx = ny = 16
    opts = CONFIGS[name]
    M, N = nx + 1, ny + 1
    orc = so.Problem(nx, ny, kkt=True, rhs_kind=1)
    s = so.Solver(orc, opts)
    ro = s.solve()
    pu, nm, ow = petsc_perm(M, N, size, 2)
    pp, _, _ = petsc_perm(M, N, size, 1)
    rng = np.random.default_rng(4)
    v = rng.uniform(-1, 1, 3 * M * N)
    yo = np.empty_like(v)
    so.lib().or_op_apply(s.ksp.contents.M, so.dptr(v), so.dptr(yo))
    vu, vp, you, yop = (np.zeros(2 * M * N), np.zeros(M * N), np.zeros(2 * M * N), np.zeros(M * N))
    vu[pu] = v[:2 * M * N]; vp[pp] = v[2 * M * N:]
    you[pu] = yo[:2 * M * N]; yop[pp] = yo[2 * M * N:]

    def rank_fn(ctx):
        prob = sp.SaddlePointProblem(ctx, nx, ny, kkt=True, rhs_kind=1)
        nl = prob.da.n_nodes_local
        g0 = int(np.sum(ow < ctx.rank))
        ksp = prob.make_ksp(opts)
        ksp.setup()
        xin = sp.Vec.from_numpy(ctx, np.concatenate([vu[2 * g0:2 * (g0 + nl)], vp[g0:g0 + nl]]))
        y = sp.Vec(ctx, prob.n)
        ksp.pc_apply(xin, y)
        yl = y.numpy()
        e = max(np.max(np.abs(yl[:2 * nl] - you[2 * g0:2 * (g0 + nl)])), np.max(np.abs(yl[2 * nl:] - yop[g0:g0 + nl])))
        x = sp.Vec(ctx, prob.n)
        r = ksp.solve(prob.rhs, x)
        rr = sp.Vec(ctx, prob.n)
        prob.K.residual(prob.rhs, x, rr)
        return e, r["its"], r["reason"], rr.norm() / prob.rhs.norm()

    res = sp.run_ranks(size, rank_fn)
    assert max(r[0] for r in res) <= 1e-10 * np.max(np.abs(yo)), [r[0] for r in res]
    assert all(r[2] == 2 for r in res)
    # LSC with unrefined CGS-GMRES is rounding-chaotic after ~20 steps (see test_gpu_parity): a few percent on the count
    tol_its = 1 if name != "fgmres_lsc" else max(1, ro["its"] // 20)
    assert abs(res[0][1] - ro["its"]) <= tol_its, (res[0][1], ro["its"])
    assert res[0][3] < 5e-7
