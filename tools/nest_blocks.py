"""Per-block SpMV time inside a nest MatMult vs each block alone (diagnostic for the spmv:C in-solve anomaly).
    python tools/nest_blocks.py [nx]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import saddle_point_petsc_b200 as sp  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 2304
ctx = sp.Context()
prob = sp.SaddlePointProblem(ctx, nx, nx, kkt=True, rhs_kind=1)
x, y = sp.Vec(ctx, prob.n), sp.Vec(ctx, prob.n)
x.set(1.0)
for _ in range(3):
    prob.K.mult(x, y)
ctx.profile(True)
for _ in range(10):
    prob.K.mult(x, y)
ctx.profile(False)
print("nest MatMult, 10 reps:", {k: round(v["ms"] / v["launches"], 4) for k, v in ctx.profile_report().items()})
for name in ("A", "Bt", "B", "C"):
    m = getattr(prob, name)
    r, c, nnz = m.size()
    xv, yv = sp.Vec(ctx, c), sp.Vec(ctx, r)
    xv.set(1.0)
    for _ in range(3):
        m.mult(xv, yv)
    ctx.timer_start()
    for _ in range(10):
        m.mult(xv, yv)
    ms = ctx.timer_stop() / 10
    ctx.timer_start()
    for _ in range(10):
        m.mult_add(xv, yv, yv)
    ms2 = ctx.timer_stop() / 10
    print("alone %s: mult %.4f ms, mult_add (z = y) %.4f ms, format %s" % (name, ms, ms2, m.spmv_format()))
