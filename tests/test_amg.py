"""Aggregation multigrid (-pc_type gamg).  The algorithm is ours (PETSc's GAMG aggregates are
not reproducible without PETSc), defined by oracle/sp_oracle_amg.c; the CPU tests pin the oracle's definition to
independent numpy restatements and to the properties an aggregation must have, the GPU tests hold the CUDA kernels to
the oracle: aggregates, node weights and tentative prolongator BIT-EXACT (integer work + one correctly rounded
division and square root), smoothed
prolongator / Galerkin operator to rounding (1e-13 relative), solves at the BASELINE tolerances (iterations +-1,
solution 1e-8)."""
import numpy as np
import pytest
import scipy.sparse as sps

import sp_oracle as so

GAMG = ("-ksp_type fgmres -ksp_gmres_restart 30 -ksp_rtol 1e-8 -pc_type fieldsplit -pc_fieldsplit_type schur "
        "-pc_fieldsplit_schur_fact_type upper -pc_fieldsplit_schur_precondition user "
        "-fieldsplit_0_ksp_type preonly -fieldsplit_0_pc_type gamg -fieldsplit_0_mg_levels_ksp_type chebyshev "
        "-fieldsplit_0_mg_levels_ksp_max_it 3 -fieldsplit_0_mg_levels_pc_type jacobi "
        "-fieldsplit_1_ksp_type preonly -fieldsplit_1_pc_type jacobi")
GAMG_PLAIN = GAMG + " -fieldsplit_0_pc_gamg_agg_nsmooths 0"
VELOCITY_GAMG = "-ksp_type gmres -ksp_rtol 1e-8 -pc_type gamg -mg_levels_ksp_max_it 2"


def key(i, order="hash"):
    if order == "natural":
        return int(i) + 1
    h = (int(i) * 2654435761) & 0xFFFFFFFF
    h ^= h >> 16
    h = (h * 0x85EBCA6B) & 0xFFFFFFFF
    h ^= h >> 13
    return ((h >> 2) << 32) | int(i)


def node_graph(A, bs):
    """Neighbour lists from block strengths sum |a| > 0 (numpy restatement of the definition)."""
    S = abs(A).tocoo()
    nn = A.shape[0] // bs
    G = sps.coo_matrix((S.data, (S.row // bs, S.col // bs)), shape=(nn, nn)).tocsr()
    G.sum_duplicates()
    G.setdiag(0)
    G.eliminate_zeros()
    return G


def aggregate_by_definition(A, bs, order="hash"):
    """Sequential restatement: repeat {largest-key undecided node with no root within distance 2 becomes a root}."""
    G = node_graph(A, bs)
    nn = G.shape[0]
    key = lambda i: globals()["key"](i, order)  # noqa: E731
    nb = [G.indices[G.indptr[i]:G.indptr[i + 1]] for i in range(nn)]
    live = [i for i in range(nn) if len(nb[i])]
    # MIS-2 by hashed priority == greedy in descending key order (a node is a root iff no higher-key root is within 2)
    root = np.zeros(nn, dtype=bool)
    blocked = np.zeros(nn, dtype=bool)
    for i in sorted(live, key=key, reverse=True):
        if blocked[i]:
            continue
        root[i] = True
        for j in nb[i]:
            blocked[j] = True
            blocked[nb[j]] = True
        blocked[i] = True
    rid = np.cumsum(root) - 1
    a1 = np.where(root, rid, -1)
    for i in live:
        if not root[i]:
            r = [j for j in nb[i] if root[j]]
            if r:
                a1[i] = rid[max(r, key=key)]
    agg = a1.copy()
    for i in live:
        if a1[i] < 0:
            r = [j for j in nb[i] if a1[j] >= 0]
            if r:
                agg[i] = a1[max(r, key=key)]
    return agg.astype(np.int32), int(root.sum())


@pytest.mark.parametrize("order", ["hash", "natural"])
@pytest.mark.parametrize("nx,ny", [(8, 8), (21, 13), (40, 40)])
def test_oracle_aggregates_follow_the_definition(nx, ny, order):
    pr = so.Problem(nx, ny, kkt=True)
    agg, nagg = so.amg_aggregate(pr.A, 2, order=order)
    ref, nref = aggregate_by_definition(pr.A.scipy(), 2, order)
    assert nagg == nref and np.array_equal(agg, ref)
    # Dirichlet nodes (identity rows) are left out, every other node belongs to exactly one aggregate
    M, N = nx + 1, ny + 1
    ii, jj = np.meshgrid(np.arange(M), np.arange(N))
    boundary = ((ii == 0) | (ii == nx) | (jj == 0) | (jj == ny)).ravel()
    assert np.all(agg[boundary] == -1) and np.all(agg[~boundary] >= 0)
    assert sorted(set(agg[agg >= 0])) == list(range(nagg))
    # scalar (pressure) block with the same routine
    aggp, naggp = so.amg_aggregate(pr.Q, 1, order=order)
    refp, nrefp = aggregate_by_definition(pr.Q.scipy(), 1, order)
    assert naggp == nrefp and np.array_equal(aggp, refp) and np.all(aggp >= 0)


def test_oracle_aggregate_sizes():
    pr = so.Problem(30, 22, kkt=True)
    G = node_graph(pr.A.scipy(), 2)
    agg, nagg = so.amg_aggregate(pr.A, 2)
    sizes = np.bincount(agg[agg >= 0], minlength=nagg)
    assert sizes.min() >= 1 and sizes.max() <= 25          # root + <=8 neighbours + second-ring joiners
    live = np.flatnonzero(np.diff(G.indptr) > 0)
    assert 6 <= len(live) / nagg <= 20                      # distance-2 independent set on a 9-point graph
    # members of one aggregate are connected through the root: any two are at most 4 edges apart
    G.data[:] = 1
    G4 = G + G @ G
    G4 = ((G4 + G4 @ G4) > 0).tocsr()
    for a in range(0, nagg, 7):
        m = np.flatnonzero(agg == a)
        sub = G4[m][:, m].toarray() | np.eye(len(m), dtype=bool)
        assert sub.all()


def test_oracle_prolongator_and_galerkin_properties():
    pr = so.Problem(24, 18, kkt=True)
    mats, interps, aggs = so.amg_hierarchy(pr.A, 2, nsmooths=1)
    assert len(mats) >= 3 and mats[-1].nrows <= 50
    agg, nagg, w0 = aggs[0]
    assert w0 is None
    Pt = so.Csr(so.lib().or_amg_tentative(len(agg), 2, so.iptr(agg), nagg, None, None)).scipy()
    assert np.allclose((Pt.T @ Pt).toarray(), np.eye(2 * nagg), atol=1e-14)       # orthonormal columns
    ones = np.zeros(pr.A.nrows); ones[0::2] = 1.0
    inside = np.repeat(agg >= 0, 2)
    assert np.allclose((Pt @ (Pt.T @ ones))[inside], ones[inside], atol=1e-14)     # constants are reproduced
    # second level: the weights are the aggregate sizes, and the two tentative prolongators together still reproduce
    # the finest level's constant (that is what carrying the weights down is for)
    agg1, nagg1, w1 = aggs[1]
    assert np.array_equal(w1, np.bincount(agg[agg >= 0], minlength=nagg))
    Pt1 = so.Csr(so.lib().or_amg_tentative(len(agg1), 2, so.iptr(agg1), nagg1, so.iptr(w1), None)).scipy()
    assert np.allclose((Pt1.T @ Pt1).toarray(), np.eye(2 * nagg1), atol=1e-14)
    c1 = Pt.T @ ones                                                                # the constant, seen from level 1
    in1 = np.repeat(agg1 >= 0, 2)
    assert np.allclose((Pt1 @ (Pt1.T @ c1))[in1], c1[in1], atol=1e-13)
    A = pr.A.scipy()
    P = interps[0].scipy()
    d = A.diagonal(); d[d == 0] = 1.0
    DA = sps.diags(1.0 / d) @ A
    lam = so.lib().or_estimate_lambda_max(so.lib().or_op_csr(pr.A.ptr), so.lib().or_op_jacobi(pr.A.ptr), 10)
    assert abs(P - (Pt - (4.0 / (3.0 * lam)) * (DA @ Pt))).max() < 1e-14
    Ac = mats[1].scipy()
    assert abs(Ac - P.T @ A @ P).max() < 1e-12 * abs(Ac).max()
    assert abs(Ac - Ac.T).max() < 1e-12 * abs(Ac).max()
    assert np.linalg.eigvalsh(mats[-1].scipy().toarray()).min() > 0


@pytest.mark.parametrize("opts", [GAMG, GAMG_PLAIN])
def test_oracle_gamg_solves_the_kkt_problem(opts):
    pr = so.Problem(24, 24, kkt=True, rhs_kind=1)
    r = so.Solver(pr, opts).solve()
    assert r["reason"] == 2
    K = pr.scipy_K()
    assert np.linalg.norm(pr.rhs - K @ r["x"]) / np.linalg.norm(pr.rhs) < 5e-8
    assert r["its"] < (40 if opts is GAMG else 90)


def test_oracle_gamg_iterations_grow_slowly_with_the_grid():
    its = []
    for nx in (16, 32, 64):
        pr = so.Problem(nx, nx, kkt=False, rhs_kind=1)
        its.append(so.Solver(pr, VELOCITY_GAMG).solve()["its"])
    assert its[-1] <= its[0] + 8 and its[-1] <= 30, its


def test_oracle_natural_ordering_gives_regular_aggregates_and_grid_independent_iterations():
    """-pc_gamg_mis_ordering natural on the lexicographically numbered grid: the roots form a lattice with spacing 3
    (aggregates of ~9 nodes instead of ~13) and the iteration count stops growing with the grid."""
    pr = so.Problem(61, 61, kkt=False)
    agg_h, nagg_h = so.amg_aggregate(pr.A, 2, order="hash")
    agg_n, nagg_n = so.amg_aggregate(pr.A, 2, order="natural")
    live = int((agg_n >= 0).sum())
    assert live == 60 * 60 and nagg_n == 20 * 20 and nagg_h < 0.8 * nagg_n
    sizes = np.bincount(agg_n[agg_n >= 0])
    assert np.count_nonzero(sizes == 9) >= 18 * 18          # all but the rows of aggregates along two sides
    its = {}
    for order in ("hash", "natural"):
        its[order] = [so.Solver(so.Problem(nx, nx, kkt=False, rhs_kind=1), VELOCITY_GAMG + " -pc_gamg_mis_ordering " + order).solve()["its"]
                      for nx in (32, 64, 128)]
    assert max(its["natural"]) - min(its["natural"]) <= 2 and its["natural"][-1] < its["hash"][-1], its


# ------------------------------------------------------------------ GPU parity
sp = None


@pytest.fixture(scope="module")
def spmod():
    global sp
    import saddle_point_petsc_b200 as m
    sp = m
    return m


def same_bits(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))


@pytest.mark.gpu
@pytest.mark.parametrize("order", ["hash", "natural"])
@pytest.mark.parametrize("nx,ny", [(3, 3), (21, 13), (64, 48), (300, 200)])
def test_aggregates_and_tentative_prolongator_bit_exact(ctx, spmod, nx, ny, order):
    dev = sp.SaddlePointProblem(ctx, nx, ny, kkt=True)
    orc = so.Problem(nx, ny, kkt=True)
    for D, O, bs in ((dev.A, orc.A, 2), (dev.Q, orc.Q, 1), (dev.C, orc.C, 1)):
        agg_d, nagg_d = D.amg_aggregate(bs, order=order)
        agg_o, nagg_o = so.amg_aggregate(O, bs, order=order)
        assert nagg_d == nagg_o and np.array_equal(agg_d, agg_o)
        if nagg_o == 0:
            continue
        rp, col, val = D.amg_prolongator(bs, order=order).csr()
        Pt = so.Csr(so.lib().or_amg_tentative(len(agg_o), bs, so.iptr(agg_o), nagg_o, None, None))
        assert np.array_equal(rp, Pt.rowptr) and np.array_equal(col, Pt.col) and same_bits(val, Pt.val)
        # with node weights (a level below the finest one)
        w = (1 + np.arange(len(agg_o)) % 17).astype(np.int32)
        wc_o = np.zeros(nagg_o, dtype=np.int32)
        Pw = so.Csr(so.lib().or_amg_tentative(len(agg_o), bs, so.iptr(agg_o), nagg_o, so.iptr(w), so.iptr(wc_o)))
        Pd, wc_d = D.amg_prolongator(bs, node_weight=w, nagg=nagg_o, order=order)
        rp, col, val = Pd.csr()
        assert np.array_equal(wc_d, wc_o)
        assert np.array_equal(rp, Pw.rowptr) and np.array_equal(col, Pw.col) and same_bits(val, Pw.val)


@pytest.mark.gpu
def test_aggregates_with_a_strength_threshold_bit_exact(ctx, spmod):
    dev = sp.SaddlePointProblem(ctx, 40, 28, kkt=True)
    orc = so.Problem(40, 28, kkt=True)
    for theta in (0.05, 0.3):
        agg_d, nagg_d = dev.A.amg_aggregate(2, theta)
        agg_o, nagg_o = so.amg_aggregate(orc.A, 2, theta)
        assert nagg_d == nagg_o and np.array_equal(agg_d, agg_o)
    assert so.amg_aggregate(orc.A, 2, 0.3)[1] != so.amg_aggregate(orc.A, 2, 0.0)[1]   # the threshold does change the graph


@pytest.mark.gpu
def test_smoothed_prolongator_and_galerkin_operator(ctx, spmod):
    dev = sp.SaddlePointProblem(ctx, 48, 36, kkt=True)
    orc = so.Problem(48, 36, kkt=True)
    omega = 0.61
    P = dev.A.amg_prolongator(2, 0.0, omega)
    agg, nagg = so.amg_aggregate(orc.A, 2)
    Pt = so.Csr(so.lib().or_amg_tentative(len(agg), 2, so.iptr(agg), nagg, None, None))
    Po = so.Csr(so.lib().or_amg_smooth_prolongator(orc.A.ptr, Pt.ptr, omega))
    rp, col, val = P.csr()
    assert np.array_equal(rp, Po.rowptr) and np.array_equal(col, Po.col)
    assert np.max(np.abs(val - Po.val)) <= 1e-13 * np.max(np.abs(Po.val))
    Ac = P.transpose().matmult(dev.A.matmult(P))
    Ao = so.Csr(so.lib().or_amg_galerkin(orc.A.ptr, Po.ptr))
    rp, col, val = Ac.csr()
    assert np.array_equal(rp, Ao.rowptr) and np.array_equal(col, Ao.col)
    assert np.max(np.abs(val - Ao.val)) <= 1e-13 * np.max(np.abs(Ao.val))


GAMG_NATURAL = GAMG + " -fieldsplit_0_pc_gamg_mis_ordering natural"


@pytest.mark.gpu
@pytest.mark.parametrize("opts", [GAMG, GAMG_PLAIN, GAMG_NATURAL])
@pytest.mark.parametrize("nx", [16, 40])
def test_gamg_pc_apply_and_solve_parity(ctx, spmod, opts, nx):
    dev = sp.SaddlePointProblem(ctx, nx, nx, kkt=True, rhs_kind=1)
    orc = so.Problem(nx, nx, kkt=True, rhs_kind=1)
    ksp = dev.make_ksp(opts)
    ksp.setup()
    s = so.Solver(orc, opts)
    v = np.random.default_rng(5).uniform(-1.0, 1.0, dev.n)
    yd = sp.Vec(ctx, dev.n)
    ksp.pc_apply(sp.Vec.from_numpy(ctx, v), yd)
    yo = np.empty(dev.n)
    so.lib().or_op_apply(s.ksp.contents.M, so.dptr(v), so.dptr(yo))
    assert np.max(np.abs(yd.numpy() - yo)) <= 1e-10 * np.max(np.abs(yo))
    x = sp.Vec(ctx, dev.n)
    rd = ksp.solve(dev.rhs, x)
    ro = s.solve()
    assert rd["reason"] == ro["reason"] == 2
    assert abs(rd["its"] - ro["its"]) <= 1, (rd["its"], ro["its"])
    xs = x.numpy()
    nu = dev.nu
    if rd["its"] == ro["its"]:
        assert np.max(np.abs(xs[:nu] - ro["x"][:nu])) <= 1e-8 * np.max(np.abs(ro["x"][:nu]))
    K = orc.scipy_K()
    assert np.linalg.norm(orc.rhs - K @ xs) / np.linalg.norm(orc.rhs) < 5e-7
    assert "smoothed aggregation" in ksp.view()


@pytest.mark.gpu
def test_gamg_on_the_velocity_block_and_on_a_scalar_matrix(ctx, spmod):
    dev = sp.SaddlePointProblem(ctx, 64, 64, kkt=False, rhs_kind=1)
    orc = so.Problem(64, 64, kkt=False, rhs_kind=1)
    ksp = dev.make_ksp(VELOCITY_GAMG)
    x = sp.Vec(ctx, dev.n)
    rd = ksp.solve(dev.rhs, x)
    ro = so.Solver(orc, VELOCITY_GAMG).solve()
    assert rd["reason"] == ro["reason"] == 2 and abs(rd["its"] - ro["its"]) <= 1, (rd["its"], ro["its"])
    if rd["its"] == ro["its"]:
        assert np.max(np.abs(x.numpy() - ro["x"])) <= 1e-8 * np.max(np.abs(ro["x"]))


@pytest.mark.gpu
def test_gamg_large_grid_properties(ctx, spmod):
    """0.5M-DOF velocity block: the aggregates keep their shape, the hierarchy is built, the solve converges."""
    dev = sp.SaddlePointProblem(ctx, 512, 512, kkt=False, rhs_kind=1)
    agg, nagg = dev.A.amg_aggregate(2)
    live = int((agg >= 0).sum())
    assert live == 511 * 511 and 6 <= live / nagg <= 20
    sizes = np.bincount(agg[agg >= 0], minlength=nagg)
    assert sizes.min() >= 1 and sizes.max() <= 25
    ksp = dev.make_ksp(VELOCITY_GAMG)
    ksp.setup()
    x = sp.Vec(ctx, dev.n)
    rd = ksp.solve(dev.rhs, x)
    assert rd["reason"] == 2 and rd["its"] <= 40
    # same solution as the geometric hierarchy's solve (both stop at rtol 1e-8 of their own preconditioned norm)
    x2 = sp.Vec(ctx, dev.n)
    r2 = dev.make_ksp("-ksp_type gmres -ksp_rtol 1e-8 -pc_type mg -pc_mg_levels 6 -mg_levels_ksp_max_it 2").solve(dev.rhs, x2)
    assert r2["reason"] == 2
    assert np.max(np.abs(x.numpy() - x2.numpy())) <= 1e-5 * np.max(np.abs(x2.numpy()))
