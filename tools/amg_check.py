"""Aggregation multigrid (-fieldsplit_0_pc_type gamg) next to the geometric hierarchy on the KKT problem:
set-up time, iterations, solve time.  python tools/amg_check.py [nx ...]"""
import json
import sys
import time

sys.path.insert(0, ".")
import bench  # noqa: E402
import saddle_point_petsc_b200 as sp  # noqa: E402

ctx = sp.Context(device=0)
for nx in [int(a) for a in sys.argv[1:]] or [576]:
    prob = sp.SaddlePointProblem(ctx, nx, nx, kkt=True, rhs_kind=1)
    base = bench.options_for("fgmres_schur_mg", nx)
    for name, extra in (("mg", ""), ("gamg", " -fieldsplit_0_pc_type gamg"),
                        ("gamg_natural", " -fieldsplit_0_pc_type gamg -fieldsplit_0_pc_gamg_mis_ordering natural"), ("gamg_plain", " -fieldsplit_0_pc_type gamg -fieldsplit_0_pc_gamg_agg_nsmooths 0")):
        ksp = prob.make_ksp(base + extra)
        x = sp.Vec(ctx, prob.n)
        t0 = time.time(); ksp.setup(); ctx.synchronize(); t_setup = time.time() - t0
        r = ksp.solve(prob.rhs, x)
        t0 = time.time(); r = ksp.solve(prob.rhs, x); ctx.synchronize(); t_solve = time.time() - t0
        rec = {"nx": nx, "dof": prob.n, "pc": name, "setup_s": round(t_setup, 3), "its": r["its"], "reason": r["reason"], "solve_ms": round(1e3 * t_solve, 2)}
        if name != "mg":
            rec["levels"] = [ln.strip() for ln in ksp.view().splitlines() if ln.strip().startswith("level ")]
        print(json.dumps(rec), flush=True)
        del ksp
