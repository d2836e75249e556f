"""BASELINE config 5: SpMV / vector-kernel bandwidth sweep, 1e5 ... 2e8 DOF, for the A, B^T, B, C blocks of the KKT
discretisation and the Krylov vector kernels, with the library's DEFAULT kernels and storage formats.

    python bench.py --config sweep [--gpus N]          (under torch.distributed.run for N > 1)
    python tools/kernel_sweep.py [nx ...]              (one GPU, explicit sizes)

Every point reports: time per launch (CUDA events, 20 launches after 3 warm-ups), the bytes the kernel has to move
(stored matrix format + x + y; vector kernels: their operand bytes) as GB/s and as a fraction of the measured HBM
peak, and -- for SpMV -- the CSR-algorithmic rate (12 nnz + 4(rows+1) + 8 rows + 8 cols).  Points whose
working set fits the 126 MB L2 are marked "l2": back-to-back launches re-read it from L2, not HBM.
For N > 1 the matrices are row-partitioned (per-rank local 32-bit indexing, halo exchange inside MatMult); the
figures are whole-job: bytes of all ranks over the slowest rank's time."""
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

L2_BYTES = 126e6
DOF_TARGETS = [1e5, 3e5, 1e6, 3e6, 1e7, 1.6e7, 3e7, 6.4e7, 1e8, 2e8]
MAX_NNZ_PER_RANK = 2.0e9       # 32-bit row pointers
MAX_DOF_PER_RANK = 1.05e8      # memory: CSR blocks + element arrays of the assembly stay well inside 180 GB


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def time_ms(ctx, fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    ctx.synchronize()
    ctx.timer_start()
    for _ in range(reps):
        fn()
    return ctx.timer_stop() / reps


def sweep_point(sp, ctx, nx, pk1, reduce_max, reduce_sum):
    pk = pk1 * max(ctx.size, 1)          # whole-job figures against the aggregate peak of the N GPUs
    prob = sp.SaddlePointProblem(ctx, nx, nx, kkt=True, rhs_kind=1)
    row = {"nx": nx, "dof": int(reduce_sum(prob.n)), "spmv": {}, "vec": {}}
    for name in ("A", "Bt", "B", "C"):
        m = getattr(prob, name)
        r, c, nnz = m.size()
        x, y = sp.Vec(ctx, c), sp.Vec(ctx, r)
        x.set(1.0)
        ms = reduce_max(time_ms(ctx, lambda: m.mult(x, y)))
        fmt = m.spmv_format()
        moved = reduce_sum(fmt["matrix_bytes"] + 8 * r + 8 * c)
        alg = reduce_sum(12 * nnz + 4 * (r + 1) + 8 * r + 8 * c)
        row["spmv"][name] = {"ms": round(ms, 5), "moved_gbs": round(moved / ms / 1e6, 1), "frac": round(moved / ms / 1e6 / pk, 4),
                             "csr_algorithmic_gbs": round(alg / ms / 1e6, 1), "bytes_per_nnz": round(fmt["matrix_bytes"] / max(nnz, 1), 2),
                             "format": "block %dx%d%s" % (fmt["block"] + (", tile dictionaries" if fmt["value_dict"] else "",)),
                             "l2": bool(moved / max(ctx.size, 1) < L2_BYTES)}
        # the same MatMult through the general (uncompressed CSR) kernel: what a matrix without repeating values gets
        m.set_spmv_format(False, False)
        ms_csr = reduce_max(time_ms(ctx, lambda: m.mult(x, y)))
        m.set_spmv_format(True, True)
        row["spmv"][name]["plain_csr_ms"] = round(ms_csr, 5)
        row["spmv"][name]["plain_csr_frac"] = round(alg / ms_csr / 1e6 / pk, 4)
        x.destroy(); y.destroy()
    n = prob.n
    a, b, w = sp.Vec(ctx, n), sp.Vec(ctx, n), sp.Vec(ctx, n)
    a.set(1.0); b.set(2.0)
    ops = (("axpy", lambda: b.axpy(1e-3, a), 24), ("waxpy", lambda: w.waxpy(0.5, a, b), 24), ("dot", lambda: a.dot(b), 16),
           ("norm", lambda: a.norm(), 8), ("copy", lambda: a.copy_to(w), 16), ("pointwise_mult", lambda: w.pointwise_mult(a, b), 24))
    for nm, fn, bpe in ops:
        ms = reduce_max(time_ms(ctx, fn))
        byts = reduce_sum(bpe * n)
        row["vec"][nm] = {"ms": round(ms, 5), "gbs": round(byts / ms / 1e6, 1), "frac": round(byts / ms / 1e6 / pk, 4), "l2": bool(bpe * n < L2_BYTES)}
    for v in (a, b, w):
        v.destroy()
    k = 15                                            # the average Gram-Schmidt step of GMRES(30)
    t_mdot, t_maxpy = ctx.bench_orthogonalization(n, k, 10)
    t_mdot, t_maxpy = reduce_max(t_mdot), reduce_max(t_maxpy)
    for nm, ms, byts in (("mdot_k15", t_mdot, 8 * (k + 1) * n), ("maxpy_norm_k15", t_maxpy, 8 * (k + 2) * n)):
        byts = reduce_sum(byts)
        row["vec"][nm] = {"ms": round(ms, 5), "gbs": round(byts / ms / 1e6, 1), "frac": round(byts / ms / 1e6 / pk, 4), "l2": False}
    for m in (prob.A, prob.Bt, prob.B, prob.C, prob.Q, prob.K):
        m.destroy()
    prob.rhs.destroy()
    prob.da.destroy()
    return row


def run(sp, ctx, sizes, dist=None):
    pk, src = peak()

    def reduce_max(v):
        if dist is None:
            return v
        import torch
        t = torch.tensor([float(v)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def reduce_sum(v):
        if dist is None:
            return v
        import torch
        t = torch.tensor([float(v)], dtype=torch.float64)
        dist.all_reduce(t)
        return float(t[0])

    rows = []
    for nx in sizes:
        rows.append(sweep_point(sp, ctx, nx, pk, reduce_max, reduce_sum))
        if ctx.rank == 0:
            sys.stderr.write(json.dumps(rows[-1]) + "\n")
            sys.stderr.flush()
    return rows, pk, src


def sizes_for(n_gpus):
    out = []
    for dof in DOF_TARGETS:
        nx = int(round(math.sqrt(dof / 3.0))) - 1
        m = nx + 1
        if 4.0 * (3 * m - 2) ** 2 / n_gpus > MAX_NNZ_PER_RANK or 3.0 * m * m / n_gpus > MAX_DOF_PER_RANK:
            continue                                   # does not fit one rank's 32-bit indexing / memory at this N
        out.append(nx)
    return out


def bench_main(args, emit):
    """bench.py --config sweep"""
    import saddle_point_petsc_b200 as sp
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist, nccl_id = None, None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("gloo", init_method="env://")
        buf = [sp.Context.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(buf, src=0)
        nccl_id = buf[0]
    ctx = sp.Context(device=local_rank, rank=rank, size=world, nccl_id=nccl_id)
    sizes = sizes_for(world)
    rows, pk, src = run(sp, ctx, sizes, dist)
    if rank != 0:
        return
    big = [r for r in rows if not r["spmv"]["A"]["l2"]]
    head = (big or rows)[-1]
    line = {"metric": "spmv_A_moved_bytes_gbs_sweep", "value": head["spmv"]["A"]["moved_gbs"], "unit": "GB/s", "n_gpus": world, "steps": 20, "warmup": 3,
            "ms_per_step": head["spmv"]["A"]["ms"], "higher_is_better": True, "scaling": "weak" if world > 1 else "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "kernel_bandwidth_sweep (BASELINE config 5)", "sizes_nx": sizes, "head_point": {"nx": head["nx"], "dof": head["dof"]},
                       "l2_policy": "no flush: points whose working set fits the 126 MB L2 are marked l2=true", "parallelism": "dmda_row_partition_x%d" % world},
            "peak": pk, "peak_source": src, "table": rows}
    emit(line)


if __name__ == "__main__":
    import saddle_point_petsc_b200 as sp
    ctx = sp.Context()
    sizes = [int(a) for a in sys.argv[1:]] or [576, 2304]
    rows, pk, src = run(sp, ctx, sizes)
    print(json.dumps({"peak_GBs": pk, "peak_source": src, "table": rows}))
