"""Multi-process (torchrun, one rank per GPU, NCCL + peer-to-peer halos) correctness check against the CPU oracle:
distributed assembly (global CSR rows), distributed MatMult of every block, and a full KKT solve.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/dist_check.py [nx]
Used by tests/test_nccl_ranks.py; the oracle is test infrastructure."""
import os
import sys

import numpy as np
import scipy.sparse as sps
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import saddle_point_petsc_b200 as sp  # noqa: E402
import sp_oracle as so  # noqa: E402


def main():
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 48
    rank, size, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("gloo", init_method="env://")
    buf = [sp.Context.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(buf, src=0)
    ctx = sp.Context(device=lr, rank=rank, size=size, nccl_id=buf[0])
    M = N = nx + 1
    orc = so.Problem(nx, nx, kkt=True, rhs_kind=1)
    nm, ow = so.dmda_natural_to_petsc(M, N, size)

    def perm(dof):
        return np.repeat(nm.astype(np.int64) * dof, dof) + np.tile(np.arange(dof), M * N)

    def petsc(A, dr, dc):
        C = A.scipy().tocoo()
        P = sps.csr_matrix((C.data, (perm(dr)[C.row], perm(dc)[C.col])), shape=C.shape)
        P.sort_indices()
        return P

    prob = sp.SaddlePointProblem(ctx, nx, nx, kkt=True, rhs_kind=1)
    nl = prob.da.n_nodes_local
    g0 = int(np.sum(ow < rank))
    rng = np.random.default_rng(3)
    xg = {1: rng.uniform(-1, 1, M * N), 2: rng.uniform(-1, 1, 2 * M * N)}
    for name, (dr, dc) in {"A": (2, 2), "Bt": (2, 1), "B": (1, 2), "C": (1, 1)}.items():
        m = getattr(prob, name)
        R = petsc(getattr(orc, name), dr, dc)
        rp, col, val = m.csr()
        Rl = R[g0 * dr:(g0 + nl) * dr]
        assert np.array_equal(rp, Rl.indptr) and np.array_equal(col, Rl.indices), name
        assert np.array_equal(val.view(np.uint64), Rl.data.view(np.uint64)), name
        for rep in range(3):                                   # repeated exchanges exercise both ghost parities
            x = sp.Vec.from_numpy(ctx, (rep + 1.0) * xg[dc][g0 * dc:(g0 + nl) * dc])
            y = sp.Vec(ctx, nl * dr)
            m.mult(x, y)
            yr = (rep + 1.0) * (R @ xg[dc])
            err = np.max(np.abs(y.numpy() - yr[g0 * dr:(g0 + nl) * dr])) / max(1.0, np.max(np.abs(yr)))
            assert err < 1e-14, (name, rep, err)
    lev = 1
    e = nx
    while e % 2 == 0 and e // 2 >= 6:
        e //= 2
        lev += 1
    opts = ("-ksp_type fgmres -ksp_rtol 1e-8 -pc_type fieldsplit -pc_fieldsplit_type schur -pc_fieldsplit_schur_fact_type upper "
            "-pc_fieldsplit_schur_precondition user -fieldsplit_0_ksp_type preonly -fieldsplit_0_pc_type mg -fieldsplit_0_pc_mg_levels %d "
            "-fieldsplit_1_ksp_type preonly -fieldsplit_1_pc_type jacobi" % max(lev, 2))
    ro = so.Solver(orc, opts).solve()
    ksp = prob.make_ksp(opts)
    x = sp.Vec(ctx, prob.n)
    for rep in range(3):                                       # the second and third solves replay the CUDA graphs
        r = ksp.solve(prob.rhs, x)
        assert r["reason"] == 2 and abs(r["its"] - ro["its"]) <= 1, (r["its"], ro["its"], r["reason"])
    xu = np.zeros(2 * M * N)
    xu[perm(2)] = ro["x"][:2 * M * N]
    eu = np.max(np.abs(x.numpy()[:2 * nl] - xu[2 * g0:2 * (g0 + nl)]))
    assert eu <= 1e-6 * np.max(np.abs(xu)), eu
    dist.barrier()
    print("rank %d ok: its %d (oracle %d)" % (rank, r["its"], ro["its"]), flush=True)


if __name__ == "__main__":
    main()
