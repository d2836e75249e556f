"""SpMV / vector-kernel bandwidth sweep (BASELINE config 5): achieved algorithmic GB/s per kernel against the
measured HBM peak.  Usage: python tools/kernel_sweep.py [nx ...]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import saddle_point_petsc_b200 as sp  # noqa: E402


def peak():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    try:
        return json.load(open(p))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


def time_ms(ctx, fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    ctx.timer_start()
    for _ in range(reps):
        fn()
    return ctx.timer_stop() / reps


def spmv_bytes(m):
    r, c, nnz = m.size()
    return 12 * nnz + 4 * (r + 1) + 8 * r + 8 * c


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [576, 2304]
    pk, src = peak()
    ctx = sp.Context()
    out = []
    for nx in sizes:
        prob = sp.SaddlePointProblem(ctx, nx, nx, kkt=True, rhs_kind=1)
        row = {"nx": nx, "dof": prob.n}
        for name in ("A", "B", "Bt", "C"):
            m = getattr(prob, name)
            r, c, nnz = m.size()
            x, y = sp.Vec(ctx, c), sp.Vec(ctx, r)
            x.set(1.0)
            for k in ([0, 1, 3] if name == "A" else [0, 3]):
                m.set_spmv_kernel(k)
                ms = time_ms(ctx, lambda: m.mult(x, y))
                gbs = spmv_bytes(m) / ms / 1e6
                row["spmv_%s_k%d" % (name, k)] = {"ms": round(ms, 4), "GBs": round(gbs, 1), "frac": round(gbs / pk, 3)}
            m.set_spmv_kernel(0)
            x.destroy(); y.destroy()
        n = prob.n
        a, b, w = sp.Vec(ctx, n), sp.Vec(ctx, n), sp.Vec(ctx, n)
        a.set(1.0); b.set(2.0)
        for nm, fn, byts in (("axpy", lambda: b.axpy(0.5, a), 24 * n), ("waxpy", lambda: w.waxpy(0.5, a, b), 24 * n),
                             ("dot", lambda: a.dot(b), 16 * n), ("norm", lambda: a.norm(), 8 * n), ("copy", lambda: a.copy_to(w), 16 * n)):
            ms = time_ms(ctx, fn)
            row[nm] = {"ms": round(ms, 4), "GBs": round(byts / ms / 1e6, 1), "frac": round(byts / ms / 1e6 / pk, 3)}
        out.append(row)
        print(json.dumps(row), flush=True)
        for v in (a, b, w):
            v.destroy()
    print(json.dumps({"peak_GBs": pk, "peak_source": src}))


if __name__ == "__main__":
    main()
