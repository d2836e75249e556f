#!/usr/bin/env python
"""bench.py -- headline benchmark: time-to-solve (rtol 1e-8) of the 2-D Stokes-type KKT system
[A B^T; B C] (16M DOF by default) with FGMRES(30) + fieldsplit-Schur on B200, plus the SpMV roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--nx NX] [--config NAME]

A "step" is one complete KSPSolve (zero initial guess, rtol 1e-8) of the assembled system.
  value   = seconds per solve with b and x resident in HBM (CUDA events on the library's compute stream)
  e2e     = the same solve through the C-ABI host-buffer entry point b200sp_ksp_solve_host: H2D copy of
            the right-hand side from pinned memory and D2H copy of the solution inside the timed region
  roofline= the A-block SpMV kernel.  `frac` is the MOVED-bytes fraction of the measured HBM peak: bytes of the
            stored (losslessly compressed) matrix format + x + y + the epilogue's operand vectors, per launch
            variant (plain / axpby / cheb), over the CUDA-event time of those launches DURING a solve.  The
            CSR-algorithmic rate (12 nnz + 4(rows+1) + 8 rows + 8 cols per launch) is kept next to it
            as `csr_algorithmic_gbs` / `speedup_vs_csr_roofline`: it exceeds 1 because the kernel does not move
            the CSR bytes.
  parity  = (N=1) the same system solved by the CPU oracle: iteration counts, final relative residuals and the
            solutions compared at north_star's tolerances (+-1, 1e-10, 1e-8); the process exits non-zero when
            they are not met.  (N>1) `dist_check`: before the timed region every rank checks, at a small size,
            its rows of every distributed matrix bit for bit against the oracle, distributed MatMult on both
            ghost-buffer parities, and one distributed solve against the oracle.
  cpu_baseline = the CPU oracle (a port of the PETSc algorithms the reference selects; PETSc itself is not
            installable here) on the host cores of the GPU box, same workload.
--impl reference times that CPU oracle as the reference arm (oracle/ is test infrastructure; this, `parity`,
`dist_check` and cpu_baseline are the only places bench.py touches it).
--config sweep runs BASELINE config 5 (SpMV / vector-kernel bandwidth sweep) instead of a solve.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS = ("-pc_type fieldsplit -pc_fieldsplit_type schur -pc_fieldsplit_schur_fact_type {fact} -pc_fieldsplit_schur_precondition user "
      "-fieldsplit_0_ksp_type preonly -fieldsplit_0_pc_type mg -fieldsplit_0_pc_mg_levels {{levels}} "
      "-fieldsplit_0_mg_levels_ksp_type chebyshev -fieldsplit_0_mg_levels_ksp_max_it 3 -fieldsplit_0_mg_levels_pc_type jacobi "
      "-fieldsplit_1_ksp_type preonly -fieldsplit_1_pc_type jacobi")
CONFIGS = {
    # BASELINE config 3: FGMRES(30), right PC, Schur factorisation with a multigrid A00 solve and the
    # pressure-mass-matrix Schur approximation
    "fgmres_schur_mg": "-ksp_type fgmres -ksp_gmres_restart 30 -ksp_rtol 1e-8 " + FS.format(fact="upper"),
    # BASELINE config 2: GMRES(30), left PC, Schur full factorisation (multigrid instead of plain Jacobi for
    # A00 so that it converges at this size)
    "gmres_schur_mg": "-ksp_type gmres -ksp_gmres_restart 30 -ksp_rtol 1e-8 " + FS.format(fact="full"),
    # BASELINE config 4 (2-D analogue): MINRES + block-diagonal, Chebyshev/Jacobi smoothed multigrid
    "minres_diag_mg": "-ksp_type minres -ksp_rtol 1e-8 " + FS.format(fact="diag"),
    # BASELINE config 3 AS NAMED: FGMRES + Schur with `self` + LSC on S (L = A10 diag(A00)^-1 A01 solved by
    # Chebyshev(8)/Jacobi).  Iterations grow with the grid: benchmarked at the size given by --nx
    "fgmres_schur_lsc": ("-ksp_type fgmres -ksp_gmres_restart 30 -ksp_rtol 1e-8 -ksp_max_it 2000 -pc_type fieldsplit -pc_fieldsplit_type schur "
                         "-pc_fieldsplit_schur_fact_type upper -pc_fieldsplit_schur_precondition self "
                         "-fieldsplit_0_ksp_type preonly -fieldsplit_0_pc_type mg -fieldsplit_0_pc_mg_levels {levels} "
                         "-fieldsplit_1_ksp_type preonly -fieldsplit_1_pc_type lsc -fieldsplit_1_pc_lsc_scale_diag "
                         "-fieldsplit_1_lsc_ksp_type chebyshev -fieldsplit_1_lsc_ksp_max_it 8 -fieldsplit_1_lsc_pc_type jacobi"),
    # BASELINE config 2 AS NAMED: GMRES + Schur (full) with plain Jacobi for A00 (iterations grow quickly with the grid)
    "gmres_schur_jacobi": ("-ksp_type gmres -ksp_gmres_restart 30 -ksp_rtol 1e-8 -ksp_max_it 20000 -pc_type fieldsplit -pc_fieldsplit_type schur "
                           "-pc_fieldsplit_schur_fact_type full -pc_fieldsplit_schur_precondition user "
                           "-fieldsplit_0_ksp_type preonly -fieldsplit_0_pc_type jacobi -fieldsplit_1_ksp_type preonly -fieldsplit_1_pc_type jacobi"),
    # config 3's solver with the A00 hierarchy built algebraically (smoothed aggregation on the device)
    # instead of from the grid: the route for matrices without a DMDA; one V-cycle per application like the default
    "fgmres_schur_gamg": ("-ksp_type fgmres -ksp_gmres_restart 30 -ksp_rtol 1e-8 -ksp_max_it 2000 -pc_type fieldsplit -pc_fieldsplit_type schur "
                          "-pc_fieldsplit_schur_fact_type upper -pc_fieldsplit_schur_precondition user "
                          "-fieldsplit_0_ksp_type preonly -fieldsplit_0_pc_type gamg -fieldsplit_0_pc_gamg_mis_ordering natural "
                          "-fieldsplit_0_mg_levels_ksp_type chebyshev -fieldsplit_0_mg_levels_ksp_max_it 3 -fieldsplit_0_mg_levels_pc_type jacobi "
                          "-fieldsplit_1_ksp_type preonly -fieldsplit_1_pc_type jacobi"),
}
# BASELINE config 4: 3-D Stokes-type KKT (Q1 hexahedra, 4 dof per node), MINRES + block-diagonal preconditioner with a
# Chebyshev/Jacobi A00 solve and the pressure-mass-matrix Schur approximation.  --nx = elements per side of the cube.
CONFIGS_3D = {
    "minres3d_diag_cheb": ("-ksp_type minres -ksp_rtol 1e-8 -ksp_max_it 20000 -pc_type fieldsplit -pc_fieldsplit_type schur -pc_fieldsplit_schur_fact_type diag "
                           "-pc_fieldsplit_schur_precondition user -fieldsplit_0_ksp_type chebyshev -fieldsplit_0_ksp_max_it 4 -fieldsplit_0_pc_type jacobi "
                           "-fieldsplit_1_ksp_type preonly -fieldsplit_1_pc_type jacobi"),
}
CONFIGS.update(CONFIGS_3D)
# north_star tolerances (BASELINE.json): iterations +-1, final relative residual 1e-10, solution rel 1e-8
TOL_ITS, TOL_RES, TOL_X = 1, 1e-10, 1e-8


def mg_levels(nx):
    """coarsen while the element count stays even and the coarse grid has >= 8 elements per side"""
    lev = 1
    while nx % 2 == 0 and nx // 2 >= 8:
        nx //= 2
        lev += 1
    return max(lev, 2)


def options_for(config, nx):
    lev = int(os.environ.get("B200SP_BENCH_MG_LEVELS", "0")) or mg_levels(nx)
    return CONFIGS[config].format(levels=lev) + (" " + os.environ["B200SP_BENCH_EXTRA_OPTS"] if os.environ.get("B200SP_BENCH_EXTRA_OPTS") else "")


def peak_hbm():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region"""

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import sp_oracle as so
    # torch.distributed.run exports OMP_NUM_THREADS=1 to every rank: the oracle takes all host threads it may use
    so.lib().or_set_threads(host_threads())
    return so


def make_problem(sp, ctx, config, nx):
    if config in CONFIGS_3D:
        return sp.SaddlePointProblem3D(ctx, nx, nx, nx, rhs_kind=1)
    return sp.SaddlePointProblem(ctx, nx, nx, kkt=True, rhs_kind=1)


def oracle_solve_setup(nx, opts, config="fgmres_schur_mg"):
    so = oracle()
    t0 = time.perf_counter()
    prob = so.Problem3D(nx, nx, nx, rhs_kind=1) if config in CONFIGS_3D else so.Problem(nx, nx, kkt=True, rhs_kind=1)
    solver = so.Solver(prob, opts)
    return so, prob, solver, time.perf_counter() - t0


def workload(args, dof, opts):
    is3d = args.config in CONFIGS_3D
    return {"workload": ("kkt3d_%s" if is3d else "kkt2d_%s") % args.config, "grid_elements": [args.nx] * (3 if is3d else 2), "dof": int(dof), "solver_options": opts,
            "rtol": 1e-8, "l2_policy": "inputs (5 GB of matrices, 128 MB vectors) far larger than the 126 MB L2; no explicit flush",
            "parallelism": "dmda_row_partition_x%d" % args.gpus}


def run_reference(args, opts):
    """reference arm: the CPU oracle on all host threads, same config.  A solve takes seconds on the CPU, so at most
    3 timed solves after at most 1 warm-up are run whatever --steps/--warmup say (the line reports what was run)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    so, prob, solver, t_setup = oracle_solve_setup(args.nx, opts, args.config)
    cores = so.lib().or_get_threads()
    steps, warmup = max(1, min(args.steps, 3)), max(0, min(args.warmup, 1))
    for _ in range(warmup):
        solver.solve(history=False)
    t0 = time.perf_counter()
    for _ in range(steps):
        r = solver.solve(history=False)
    dt = (time.perf_counter() - t0) / steps
    line = {"impl": "reference", "metric": "time_to_solve_rtol1e-8", "value": dt, "unit": "s", "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "steps_requested": args.steps, "warmup_requested": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload(args, prob.nu + prob.np_, opts),
            "iterations": r["its"], "converged_reason": r["reason"],
            "cpu_baseline": {"value": dt, "unit": "s", "cores": cores, "kind": "port",
                             "sample": "full workload: one complete solve per step, %d timed (assembly+setup %.1fs untimed)" % (steps, t_setup)},
            "setup_s": t_setup,
            "e2e": {"value": dt, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


_REAL_STDOUT = None


def emit(line):
    """the ONE JSON line of the contract, on the process's real stdout"""
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


def compare_with_oracle(x, ro, rd, nu, perm_u=None, perm_p=None, g0=0, nl=None):
    """GPU solution (this rank's part) against the oracle's: the quantities north_star names.  perm_*: natural ->
    PETSc numbering when the GPU run is row-partitioned."""
    xo = ro["x"]
    nug = len(perm_u) if perm_u is not None else nu
    xu, xp = xo[:nug], xo[nug:]
    if perm_u is not None:
        t = np.zeros(nug); t[perm_u] = xu; xu = t[2 * g0:2 * (g0 + nl)]
        t = np.zeros(len(xp)); t[perm_p] = xp; xp = t[g0:g0 + nl]
    du = float(np.max(np.abs(x[:nu] - xu)))
    dp = x[nu:] - xp
    return {"du": du, "umax": float(np.max(np.abs(ro["x"][:nug]))), "dp": dp, "pmax": float(np.max(np.abs(ro["x"][nug:])))}


def dist_check(sp, ctx, dist, nx=96):
    """tools/dist_check.py in-process (N>1, before the timed region): global CSR rows bit-exact against the oracle,
    distributed MatMult on both ghost parities, one distributed solve against the oracle."""
    import scipy.sparse as sps
    so = oracle()
    rank, size = ctx.rank, ctx.size
    M = N = nx + 1
    out = {"nx": nx, "ok": False}
    try:
        orc = so.Problem(nx, nx, kkt=True, rhs_kind=1)
        nm, ow = so.dmda_natural_to_petsc(M, N, size)

        def perm(dof):
            return np.repeat(nm.astype(np.int64) * dof, dof) + np.tile(np.arange(dof), M * N)

        def petsc(A, dr, dc):
            Cm = A.scipy().tocoo()
            P = sps.csr_matrix((Cm.data, (perm(dr)[Cm.row], perm(dc)[Cm.col])), shape=Cm.shape)
            P.sort_indices()
            return P

        prob = sp.SaddlePointProblem(ctx, nx, nx, kkt=True, rhs_kind=1)
        nl = prob.da.n_nodes_local
        g0 = int(np.sum(ow < rank))
        rng = np.random.default_rng(3)
        xg = {1: rng.uniform(-1, 1, M * N), 2: rng.uniform(-1, 1, 2 * M * N)}
        bit_exact, mm_err = True, 0.0
        for name, (dr, dc) in {"A": (2, 2), "Bt": (2, 1), "B": (1, 2), "C": (1, 1)}.items():
            m = getattr(prob, name)
            R = petsc(getattr(orc, name), dr, dc)
            rp, col, val = m.csr()
            Rl = R[g0 * dr:(g0 + nl) * dr]
            bit_exact = bit_exact and np.array_equal(rp, Rl.indptr) and np.array_equal(col, Rl.indices) and \
                np.array_equal(val.view(np.uint64), Rl.data.view(np.uint64))
            for rep in range(3):                                   # repeated exchanges exercise both ghost parities
                x = sp.Vec.from_numpy(ctx, (rep + 1.0) * xg[dc][g0 * dc:(g0 + nl) * dc])
                y = sp.Vec(ctx, nl * dr)
                m.mult(x, y)
                yr = (rep + 1.0) * (R @ xg[dc])
                mm_err = max(mm_err, float(np.max(np.abs(y.numpy() - yr[g0 * dr:(g0 + nl) * dr])) / max(1.0, np.max(np.abs(yr)))))
        opts = CONFIGS["fgmres_schur_mg"].format(levels=mg_levels(nx))
        ro = so.Solver(orc, opts).solve()
        ksp = prob.make_ksp(opts)
        x = sp.Vec(ctx, prob.n)
        its = []
        for rep in range(3):                                       # the second and third solves replay the CUDA graphs
            r = ksp.solve(prob.rhs, x)
            its.append(r["its"])
        c = compare_with_oracle(x.numpy(), ro, r, 2 * nl, perm(2), perm(1), g0, nl)
        loc = {"bit_exact": bool(bit_exact), "mm_err": mm_err, "its": its, "reason": r["reason"], "du": c["du"], "dp": c["dp"],
               "rel_res": r["rnorm"] / r["history"][0]}
        ksp.destroy()
    except Exception as e:  # noqa: BLE001 -- reported in the JSON line; the bench goes on so that the failure is visible
        loc = {"error": repr(e)}
        ro = None
    allr = [None] * size
    dist.all_gather_object(allr, loc)
    if rank != 0:
        return None
    errs = [a["error"] for a in allr if "error" in a]
    if errs or ro is None:
        out["error"] = errs[0] if errs else "oracle failed"
        return out
    umax, pmax = float(np.max(np.abs(ro["x"][:2 * M * N]))), float(np.max(np.abs(ro["x"][2 * M * N:])))
    dp = np.concatenate([a["dp"] for a in allr])
    rel_o = ro["rnorm"] / ro["history"][0]
    out.update({"csr_rows_bit_exact": all(a["bit_exact"] for a in allr), "matmult_max_rel_err": max(a["mm_err"] for a in allr),
                "solve_iterations": allr[0]["its"], "oracle_iterations": ro["its"],
                "rel_residual_diff": abs(allr[0]["rel_res"] - rel_o),
                "max_rel_velocity_diff": max(a["du"] for a in allr) / umax,
                "max_rel_pressure_diff_mean_removed": float(np.max(np.abs(dp - dp.mean()))) / pmax})
    out["ok"] = bool(out["csr_rows_bit_exact"] and out["matmult_max_rel_err"] < 1e-14 and all(a["reason"] == 2 for a in allr)
                     and all(abs(i - ro["its"]) <= TOL_ITS for a in allr for i in a["its"])
                     and out["max_rel_velocity_diff"] <= TOL_X and out["max_rel_pressure_diff_mean_removed"] <= TOL_X
                     and (allr[0]["its"][-1] != ro["its"] or out["rel_residual_diff"] <= TOL_RES))
    return out


def time_solves(ctx, ksp, b, x, steps, warmup):
    for _ in range(warmup):
        ksp.solve(b, x)
    ctx.synchronize()
    ctx.timer_start()
    for _ in range(steps):
        res = ksp.solve(b, x)
    ms = ctx.timer_stop()
    return ms / 1e3 / steps, res


def secondary_config(sp, ctx, name, nx, steps=3, warmup=2):
    """one more BASELINE configuration on this GPU (N=1 only), device-resident timing, true residual reported"""
    opts = options_for(name, nx)
    prob = sp.SaddlePointProblem(ctx, nx, nx, kkt=True, rhs_kind=1)
    t0 = time.perf_counter()
    ksp = prob.make_ksp(opts)
    ksp.setup()
    ctx.synchronize()
    t_setup = time.perf_counter() - t0
    x = sp.Vec(ctx, prob.n)
    dt, res = time_solves(ctx, ksp, prob.rhs, x, steps, warmup)
    r = sp.Vec(ctx, prob.n)
    prob.K.residual(prob.rhs, x, r)
    out = {"config": name, "grid_elements": [nx, nx], "dof": prob.n, "solver_options": opts, "value": dt, "unit": "s", "steps": steps,
           "warmup": warmup, "iterations": res["its"], "converged_reason": res["reason"], "true_relative_residual": r.norm() / prob.rhs.norm(),
           "ksp_setup_s": t_setup}
    ksp.destroy()
    return out


def spmv_moved_bytes(fmt, rows, cols, variant):
    """bytes one SpMV launch has to move: stored matrix format + x + y + the epilogue's operand vectors"""
    extra = {"plain": 0, "axpby": 8 * rows, "cheb": 32 * rows}[variant]   # z | z, dinv, p_{k-1}, p_k
    return fmt["matrix_bytes"] + 8 * rows + 8 * cols + extra


def main():
    # Libraries (NCCL's version banner, torch.distributed notices) write to fd 1; the contract is exactly one JSON
    # line on stdout.  Everything else goes to stderr: fd 1 is pointed at fd 2 and the result is written to the
    # saved descriptor at the end.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nx", type=int, default=2304, help="elements per side (2304 -> 15.9M DOF; 576 -> 1.0M DOF)")
    ap.add_argument("--config", default="fgmres_schur_mg", choices=sorted(CONFIGS) + ["sweep"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU oracle legs (cpu_baseline and parity)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary configurations (N=1)")
    ap.add_argument("--no-dist-check", action="store_true", help="skip the in-run distributed parity check (N>1)")
    args = ap.parse_args()

    if args.config == "sweep":
        import tools.kernel_sweep as ks
        ks.bench_main(args, emit)
        return
    opts = options_for(args.config, args.nx)
    if args.impl == "reference":
        run_reference(args, opts)
        return

    import saddle_point_petsc_b200 as sp
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("bench.py: --gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run for N>1)" % (args.gpus, world))
    dist = None
    nccl_id = None
    if world > 1:
        # one process per GPU: torch.distributed (gloo, 127.0.0.1) is only the bootstrap for the NCCL unique id and
        # the max-over-ranks of the timings; the solver's collectives are NCCL calls inside libb200sp
        # rank 0 must print exactly ONE line on stdout: keep NCCL's own banner ("NCCL version ...", printed when
        # NCCL_DEBUG=VERSION/INFO) off stdout
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO", "TRACE"):
            os.environ["NCCL_DEBUG_FILE"] = os.environ.get("NCCL_DEBUG_FILE", "/tmp/b200sp_nccl_%h_%p.log")
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("gloo", init_method="env://")
        buf = [sp.Context.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(buf, src=0)
        nccl_id = buf[0]

    def max_over_ranks(v):
        if dist is None:
            return v
        import torch
        t = torch.tensor([float(v)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def barrier():
        if dist is not None:
            dist.barrier()

    ctx = sp.Context(device=local_rank, rank=rank, size=world, nccl_id=nccl_id)
    dcheck = None
    if world > 1 and not args.no_dist_check and args.config not in CONFIGS_3D:
        dcheck = dist_check(sp, ctx, dist)
        barrier()
    # A fresh box starts cold (lazy kernel loading, clocks, allocator): run the whole path once on a 1M-DOF grid so that
    # `assembly_s` and `ksp_setup_s` below measure the work (measured on a fresh box: 0.32 s cold vs 0.08 s warm)
    # (N > 1: dist_check above has already run every kernel)
    if args.config not in CONFIGS_3D and world == 1:
        wnx = 576 if args.nx >= 576 else 32
        wp = sp.SaddlePointProblem(ctx, wnx, wnx, kkt=True, rhs_kind=1)
        wk = wp.make_ksp(options_for(args.config, wnx))
        for _ in range(3):
            wk.solve(wp.rhs, sp.Vec(ctx, wp.n))
        wk.destroy()
        del wp, wk
    ctx.synchronize()
    barrier()
    t0 = time.perf_counter()
    prob = make_problem(sp, ctx, args.config, args.nx)
    ctx.synchronize()
    t_assembly = time.perf_counter() - t0
    t0 = time.perf_counter()
    ksp = prob.make_ksp(opts)
    ksp.setup()
    ctx.synchronize()
    t_setup = time.perf_counter() - t0
    n = prob.n
    x = sp.Vec(ctx, n)

    # ---- device-resident timing: W warm-up solves, then exactly K timed solves between events + syncs
    t0 = time.perf_counter()
    ksp.solve(prob.rhs, x)     # the first solve also builds the lazily derived SpMV formats and records the CUDA graphs
    ctx.synchronize()
    t_first_solve = time.perf_counter() - t0
    for _ in range(max(args.warmup - 1, 0)):
        ksp.solve(prob.rhs, x)
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    ctx.synchronize()
    barrier()
    l0 = ctx.launch_count()
    ctx.timer_start()
    for _ in range(args.steps):
        res = ksp.solve(prob.rhs, x)
    ms = ctx.timer_stop()
    ctx.synchronize()
    barrier()
    launches = ctx.launch_count() - l0
    clocks = sampler.stop()
    t_solve = max_over_ranks(ms) / 1e3 / args.steps   # device time, max over ranks
    x_dev = x.numpy() if (world == 1 and not args.no_cpu_baseline) else None

    # ---- end to end through the host-buffer C-ABI entry point, pinned host memory
    import torch
    hb = torch.empty(n, dtype=torch.float64).pin_memory()
    hx = torch.empty(n, dtype=torch.float64).pin_memory()
    b_np, x_np = hb.numpy(), hx.numpy()
    b_np[:] = prob.rhs.numpy()
    ksp.solve_host(b_np, x_np)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ksp.solve_host(b_np, x_np)
    t_e2e = max_over_ranks((time.perf_counter() - t0) / args.steps)

    # ---- per-kernel-class device time during one solve (events around every launch; measurement pass only)
    barrier()
    ctx.profile(True)
    ksp.solve(prob.rhs, x)
    ctx.profile(False)
    barrier()
    prof = ctx.profile_report()
    pk, pk_src = peak_hbm()
    rA, cA, nnzA = prob.A.size()
    bytes_A = 12 * nnzA + 4 * (rA + 1) + 8 * rA + 8 * cA
    fmt = prob.A.spmv_format()
    variants, tot_ms, tot_n, tot_moved = {}, 0.0, 0, 0.0
    for key, v in prof.items():
        if not key.startswith("spmv:A|") or not v["launches"]:
            continue
        var = key.split("|")[1]
        moved = spmv_moved_bytes(fmt, rA, cA, var)
        avg = v["ms"] / v["launches"]
        variants[var] = {"launches_per_solve": v["launches"], "avg_launch_ms": round(avg, 5), "moved_bytes_per_launch": moved,
                         "moved_gbs": round(moved / avg / 1e6, 1), "frac": round(moved / avg / 1e6 / pk, 4),
                         "csr_algorithmic_gbs": round(bytes_A / avg / 1e6, 1)}
        tot_ms += v["ms"]; tot_n += v["launches"]; tot_moved += moved * v["launches"]
    avg_ms = tot_ms / max(tot_n, 1)
    moved_avg = tot_moved / max(tot_n, 1)
    total_prof = sum(v["ms"] for v in prof.values())
    traffic = None   # dram__bytes_read.sum + dram__bytes_write.sum of the plain-epilogue launch of this kernel on this matrix (ncu --set full)
    for tf in ("r02_spmv_traffic.json",):
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", tf)))
            if tr["nx"] == args.nx and tr["n_gpus"] == world and tr.get("format") == fmt.get("format", tr.get("format")) and \
                    tr.get("value_dict", False) == fmt["value_dict"]:
                traffic = tr["traffic"]
                break
        except Exception:
            pass
    kname = fmt.get("kernel") or (("k_spmv_pd<%d,%d,3> (tile-local pattern/value dictionaries)" % fmt["block"]) if fmt["value_dict"] else
                                  ("k_spmv_tma_blk<%d,%d>" % fmt["block"]) if fmt["block"] != (1, 1) else "k_spmv_tma")
    achieved = moved_avg / avg_ms / 1e6 if avg_ms > 0 else 0.0
    roofline = {"kernel": "%s on the A block (%d x %d, %d nnz per GPU)" % (kname, rA, cA, nnzA), "bound": "hbm",
                "achieved": round(achieved, 1), "peak": pk, "peak_source": pk_src, "unit": "GB/s", "frac": round(achieved / pk, 4),
                "frac_definition": "bytes the kernel has to move (stored matrix format + x + y + epilogue operand vectors, launch-weighted over "
                                   "the variants below) / CUDA-event time of the A-block launches inside one solve / measured HBM peak",
                "traffic": traffic, "traffic_note": "ncu dram bytes of ONE plain-epilogue launch (profiles/); compare with variants.plain.moved_bytes_per_launch",
                "moved_bytes_per_launch_avg": int(moved_avg), "stored_matrix_bytes": fmt["matrix_bytes"],
                "csr_algorithmic_bytes_per_launch": bytes_A, "csr_algorithmic_gbs": round(bytes_A / avg_ms / 1e6, 1) if avg_ms > 0 else None,
                "speedup_vs_csr_roofline": round(bytes_A / avg_ms / 1e6 / pk, 4) if avg_ms > 0 else None,
                "avg_launch_ms": round(avg_ms, 5), "launches_per_solve": tot_n, "variants": variants,
                "share_of_solve_device_time": round(tot_ms / total_prof, 4) if total_prof else None}
    # classes: epilogue variants of one matrix merged back; "profiled": every launch was bracketed by events and a
    # host synchronisation, so the SUM exceeds `value` -- use the shares, not the absolute numbers
    merged = {}
    for k, v in prof.items():
        kk = k.split("|")[0]
        e = merged.setdefault(kk, {"ms": 0.0, "launches": 0})
        e["ms"] += v["ms"]; e["launches"] += v["launches"]
    classes = {k: {"ms": round(v["ms"], 3), "launches": v["launches"]} for k, v in sorted(merged.items(), key=lambda kv: -kv[1]["ms"])}

    # true residual of the last solve (device SpMV), reported for the record
    r = sp.Vec(ctx, n)
    prob.K.residual(prob.rhs, x, r)
    true_rel = r.norm() / prob.rhs.norm()

    if dist is not None:
        t = torch.tensor([float(n)], dtype=torch.float64)
        dist.all_reduce(t)
        n_global = int(t[0])
        t = torch.tensor([float(launches)], dtype=torch.float64)
        dist.all_reduce(t)
        launches = int(t[0])
    else:
        n_global = n
    if rank != 0:
        ctx.synchronize()
        return

    # ---- N=1: the same system through the CPU oracle -> cpu_baseline + parity at the benchmarked size
    cpu, parity = None, None
    if not args.no_cpu_baseline and world == 1:
        so, oprob, osolver, t_osetup = oracle_solve_setup(args.nx, opts, args.config)
        t0 = time.perf_counter()
        orr = osolver.solve(history=True)
        t_cpu = time.perf_counter() - t0
        cpu = {"value": t_cpu, "unit": "s", "cores": so.lib().or_get_threads(), "kind": "port",
               "sample": "full workload, one solve (oracle assembly+setup %.1fs untimed)" % t_osetup, "iterations": orr["its"],
               "setup_s": t_osetup}
        c = compare_with_oracle(x_dev, orr, res, prob.nu)
        rel_d, rel_o = res["rnorm"] / res["history"][0], orr["rnorm"] / orr["history"][0]
        m = min(len(res["history"]), len(orr["history"]))
        hist_dev = float(np.max(np.abs(res["history"][:m] - orr["history"][:m]) / orr["history"][:m])) if m else None
        parity = {"against": "CPU oracle (oracle/sp_oracle.c), same grid, right-hand side and options", "size": "nx=%d (benchmarked size)" % args.nx,
                  "iterations": res["its"], "oracle_iterations": orr["its"], "reason": res["reason"], "oracle_reason": orr["reason"],
                  "rel_residual": rel_d, "oracle_rel_residual": rel_o, "rel_residual_diff": abs(rel_d - rel_o),
                  "max_rel_velocity_diff": c["du"] / c["umax"],
                  "max_rel_pressure_diff_mean_removed": float(np.max(np.abs(c["dp"] - c["dp"].mean()))) / c["pmax"],
                  "max_rel_history_diff": hist_dev,
                  "tolerances": {"iterations": TOL_ITS, "rel_residual": TOL_RES, "solution": TOL_X}}
        tol_its, tol_x = TOL_ITS, TOL_X
        if args.config in ("fgmres_schur_lsc", "gmres_schur_jacobi", "fgmres_schur_gamg"):
            # hundreds of unrefined classical-Gram-Schmidt steps on a weak preconditioner: the iteration COUNT itself moves by a
            # few percent under any change of summation order (the oracle moves by as much under a mathematically neutral
            # rescaling, tests/test_gpu_parity.py), and two valid rtol-1e-8 iterates then differ by about the last correction
            tol_its, tol_x = max(1, orr["its"] // 20), 1e-6
            parity["tolerances"] = {"iterations": tol_its, "rel_residual": TOL_RES, "solution": tol_x,
                                    "note": "weakly preconditioned configuration: tolerances are the oracle's own measured sensitivity"}
        parity["ok"] = bool(res["reason"] == orr["reason"] and abs(res["its"] - orr["its"]) <= tol_its
                            and (res["its"] != orr["its"] or parity["rel_residual_diff"] <= TOL_RES)
                            and parity["max_rel_velocity_diff"] <= tol_x and parity["max_rel_pressure_diff_mean_removed"] <= tol_x)
        del osolver, oprob

    # ---- the A-block MatMult in every storage format the library has (N=1, after the last solve: switching formats drops
    # the derived storage the recorded CUDA graphs point at), and on a variable-coefficient operator, whose values do not
    # repeat: what a general matrix gets.  Bare launches, CUDA events, 20 after 3 warm-ups; results are bit-identical.
    if world == 1 and args.config not in CONFIGS_3D and not args.no_secondary:
        def time_mult(m, rows, cols):
            xv, yv = sp.Vec(ctx, cols), sp.Vec(ctx, rows)
            xv.set(1.0)
            for _ in range(3):
                m.mult(xv, yv)
            ctx.synchronize()
            ctx.timer_start()
            for _ in range(20):
                m.mult(xv, yv)
            t = ctx.timer_stop() / 20
            f = m.spmv_format()
            moved = f["matrix_bytes"] + 8 * rows + 8 * cols
            xv.destroy(); yv.destroy()
            return {"ms": round(t, 5), "block": list(f["block"]), "value_dict": f["value_dict"], "moved_bytes": moved,
                    "moved_gbs": round(moved / t / 1e6, 1), "frac": round(moved / t / 1e6 / pk, 4),
                    "csr_algorithmic_gbs": round(bytes_A / t / 1e6, 1)}
        formats = {"tile dictionaries (default)": time_mult(prob.A, rA, cA)}
        prob.A.set_spmv_format(True, False)
        formats["block column index + plain values"] = time_mult(prob.A, rA, cA)
        prob.A.set_spmv_format(False, False)
        formats["plain CSR"] = time_mult(prob.A, rA, cA)
        prob.A.set_spmv_format(True, True)
        try:
            Av = prob.da.assemble_stress_coeff(1)
            formats["variable-coefficient operator, default policy"] = time_mult(Av, rA, cA)
            Av.destroy()
        except Exception as e:  # noqa: BLE001
            formats["variable-coefficient operator, default policy"] = {"error": repr(e)}
        roofline["formats_bare_matmult"] = formats

    secondary = None
    if world == 1 and not args.no_secondary and args.config == "fgmres_schur_mg":
        secondary = []
        # config 2 with multigrid (1M DOF) and AS NAMED with plain Jacobi for A00 (at the size where it still converges),
        # config 3 AS NAMED (LSC, largest size converging in <= 500 iterations), config 4's solver in 2-D at the full size
        for name, nx2 in (("gmres_schur_mg", 576), ("gmres_schur_jacobi", 64), ("fgmres_schur_lsc", 96), ("fgmres_schur_gamg", 576),
                          ("minres_diag_mg", args.nx)):
            try:
                secondary.append(secondary_config(sp, ctx, name, nx2))
            except Exception as e:  # noqa: BLE001
                secondary.append({"config": name, "grid_elements": [nx2, nx2], "error": repr(e)})

    cfg = workload(args, n_global, opts)   # identical to the reference arm's `config` (same keys, same values)
    line = {"metric": "time_to_solve_rtol1e-8", "value": t_solve, "unit": "s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_solve * 1e3, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": cfg, "dof_per_gpu_rank0": n, "iterations": res["its"], "converged_reason": res["reason"],
            "iterations_per_s": res["its"] / t_solve, "true_relative_residual": true_rel,
            "e2e": {"value": t_e2e, "unit": "s", "h2d_bytes_per_step": 8 * n_global, "d2h_bytes_per_step": 8 * n_global},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "dist_check": dcheck,
            "kernel_classes_ms_per_solve": {"profiled": True, "note": "one extra solve with CUDA events + a host sync around every launch: use shares",
                                            "total_ms": round(total_prof, 3), "classes": classes},
            "assembly_s": t_assembly, "ksp_setup_s": t_setup, "first_solve_s": t_first_solve,
            "setup_plus_solve_s": t_setup + t_first_solve, "secondary": secondary}
    emit(line)
    bad = (parity is not None and not parity["ok"]) or (dcheck is not None and not dcheck["ok"])
    if bad:
        sys.stderr.write("bench.py: PARITY CHECK FAILED: %s\n" % json.dumps(parity if parity is not None and not parity["ok"] else dcheck))
        sys.stderr.flush()
        os._exit(3)


if __name__ == "__main__":
    main()
