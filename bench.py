#!/usr/bin/env python
"""bench.py -- headline benchmark: time-to-solve (rtol 1e-8) of the 2-D Stokes-type KKT system
[A B^T; B C] (16M DOF by default) with FGMRES(30) + fieldsplit-Schur on B200, plus the SpMV roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--nx NX] [--config NAME]

A "step" is one complete KSPSolve (zero initial guess, rtol 1e-8) of the assembled system.
  value   = seconds per solve with b and x resident in HBM (CUDA events on the library's compute stream)
  e2e     = the same solve through the C-ABI host-buffer entry point b200sp_ksp_solve_host: H2D copy of
            the right-hand side from pinned memory and D2H copy of the solution inside the timed region
  roofline= the A-block SpMV kernel: algorithmic bytes (12 nnz + 4(rows+1) + 8 rows + 8 cols) / its average
            launch duration measured with CUDA events DURING a solve, against MEASURED_PEAKS.json
  cpu_baseline = the CPU oracle (a port of the PETSc algorithms the reference selects; PETSc itself is not
            installable here) on the host cores of the GPU box, same workload.
--impl reference times that CPU oracle as the reference arm (oracle/ is test infrastructure; this and
cpu_baseline are the only places bench.py touches it).
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

CONFIGS = {
    # BASELINE config 3: FGMRES(30), right PC, Schur factorisation with a multigrid A00 solve and the
    # pressure-mass-matrix Schur approximation
    "fgmres_schur_mg": ("-ksp_type fgmres -ksp_gmres_restart 30 -ksp_rtol 1e-8 -pc_type fieldsplit -pc_fieldsplit_type schur "
                        "-pc_fieldsplit_schur_fact_type upper -pc_fieldsplit_schur_precondition user "
                        "-fieldsplit_0_ksp_type preonly -fieldsplit_0_pc_type mg -fieldsplit_0_pc_mg_levels {levels} "
                        "-fieldsplit_1_ksp_type preonly -fieldsplit_1_pc_type jacobi"),
    # BASELINE config 2: GMRES(30), left PC, Schur full factorisation (multigrid instead of plain Jacobi for
    # A00 so that it converges at this size)
    "gmres_schur_mg": ("-ksp_type gmres -ksp_gmres_restart 30 -ksp_rtol 1e-8 -pc_type fieldsplit -pc_fieldsplit_type schur "
                       "-pc_fieldsplit_schur_fact_type full -pc_fieldsplit_schur_precondition user "
                       "-fieldsplit_0_ksp_type preonly -fieldsplit_0_pc_type mg -fieldsplit_0_pc_mg_levels {levels} "
                       "-fieldsplit_1_ksp_type preonly -fieldsplit_1_pc_type jacobi"),
    # BASELINE config 4 (2-D analogue): MINRES + block-diagonal, Chebyshev/Jacobi smoothed multigrid
    "minres_diag_mg": ("-ksp_type minres -ksp_rtol 1e-8 -pc_type fieldsplit -pc_fieldsplit_type schur "
                       "-pc_fieldsplit_schur_fact_type diag -pc_fieldsplit_schur_precondition user "
                       "-fieldsplit_0_ksp_type preonly -fieldsplit_0_pc_type mg -fieldsplit_0_pc_mg_levels {levels} "
                       "-fieldsplit_1_ksp_type preonly -fieldsplit_1_pc_type jacobi"),
}


def mg_levels(nx):
    """coarsen while the element count stays even and the coarse grid has >= 8 elements per side"""
    lev = 1
    while nx % 2 == 0 and nx // 2 >= 8:
        nx //= 2
        lev += 1
    return max(lev, 2)


def peak_hbm():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


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


def oracle_solve_setup(nx, opts):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import sp_oracle as so
    t0 = time.perf_counter()
    prob = so.Problem(nx, nx, kkt=True, rhs_kind=1)
    solver = so.Solver(prob, opts)
    return so, prob, solver, time.perf_counter() - t0


def run_reference(args, opts):
    """reference arm: the CPU oracle on all host threads, same config; K timed solves after W warm-ups"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    so, prob, solver, t_setup = oracle_solve_setup(args.nx, opts)
    cores = so.lib().or_get_threads()
    for _ in range(max(args.warmup, 0)):
        solver.solve(history=False)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = solver.solve(history=False)
    dt = (time.perf_counter() - t0) / args.steps
    line = {"impl": "reference", "metric": "time_to_solve_rtol1e-8", "value": dt, "unit": "s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload(args, prob.nu + prob.np_),
            "iterations": r["its"], "converged_reason": r["reason"],
            "cpu_baseline": {"value": dt, "unit": "s", "cores": cores, "kind": "port",
                             "sample": "full workload: one complete solve per step (assembly %.1fs and setup untimed)" % t_setup},
            "e2e": {"value": dt, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def workload(args, dof):
    return {"workload": "kkt2d_%s" % args.config, "grid_elements": [args.nx, args.nx], "dof": int(dof), "solver_options": None,
            "rtol": 1e-8, "l2_policy": "inputs (5 GB of matrices, 128 MB vectors) far larger than the 126 MB L2; no explicit flush",
            "parallelism": "dmda_row_partition_x%d" % args.gpus}


_REAL_STDOUT = None


def emit(line):
    """the ONE JSON line of the contract, on the process's real stdout"""
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


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
    ap.add_argument("--config", default="fgmres_schur_mg", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    ap_levels = int(os.environ.get("B200SP_BENCH_MG_LEVELS", "0")) or mg_levels(args.nx)
    opts = CONFIGS[args.config].format(levels=ap_levels)

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
    t0 = time.perf_counter()
    prob = sp.SaddlePointProblem(ctx, args.nx, args.nx, kkt=True, rhs_kind=1)
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
    for _ in range(args.warmup):
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
    pa = prof.get("spmv:A", {"ms": 0.0, "launches": 0})
    avg_ms = pa["ms"] / max(pa["launches"], 1)
    achieved = bytes_A / avg_ms / 1e6 if avg_ms > 0 else 0.0
    # wall-clock split of one un-profiled solve is not available per class; report the host-side time of the same
    # profiled solve next to the sum of its device classes so launch gaps / exposed communication are visible
    prof_total_ms = sum(v["ms"] for v in prof.values())
    total_prof = sum(v["ms"] for v in prof.values())
    traffic = None   # dram__bytes_read.sum + dram__bytes_write.sum of this kernel on this matrix, from the committed ncu capture
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_spmv_traffic.json")))
        if tr["nx"] == args.nx and tr["n_gpus"] == world and tr.get("value_dict", False) == prob.A.spmv_format()["value_dict"]:
            traffic = tr["traffic"]
    except Exception:
        pass
    fmt = prob.A.spmv_format()
    kname = ("k_spmv_tma_dict<%d,%d>" % fmt["block"]) if fmt["value_dict"] else ("k_spmv_tma_blk<%d,%d>" % fmt["block"]) if fmt["block"] != (1, 1) else "k_spmv_tma"
    stored = fmt["matrix_bytes"] + 8 * rA + 8 * cA   # what the kernel has to move for the stored (losslessly compressed) format
    # `achieved` follows the contract: ALGORITHMIC (plain CSR) bytes / launch time.  The kernel streams a
    # compressed matrix (block column index + tile-local value dictionary), so this can exceed the HBM peak; the
    # bytes really moved are `traffic` (ncu) ~ `stored_format_bytes_per_launch`, and `frac_of_peak_moved` is that rate.
    roofline = {"kernel": "%s on the A block (%d x %d, %d nnz per GPU)" % (kname, rA, cA, nnzA), "bound": "hbm", "achieved": round(achieved, 1),
                "peak": pk, "peak_source": pk_src, "unit": "GB/s", "frac": round(achieved / pk, 4), "traffic": traffic,
                "algorithmic_bytes_per_launch": bytes_A, "stored_format_bytes_per_launch": stored,
                "frac_of_peak_moved": round((traffic if traffic else stored) / avg_ms / 1e6 / pk, 4) if avg_ms > 0 else None,
                "avg_launch_ms": round(avg_ms, 5), "launches_per_solve": pa["launches"],
                "share_of_solve_device_time": round(pa["ms"] / total_prof, 4) if total_prof else None}
    classes = {k: {"ms": round(v["ms"], 3), "launches": v["launches"]} for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}

    # true residual of the last solve (device SpMV), reported for the record
    r = sp.Vec(ctx, n)
    prob.K.residual(prob.rhs, x, r)
    true_rel = r.norm() / prob.rhs.norm()

    if dist is not None:
        import torch
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
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        so, oprob, osolver, t_osetup = oracle_solve_setup(args.nx, opts)
        t0 = time.perf_counter()
        orr = osolver.solve(history=False)
        t_cpu = time.perf_counter() - t0
        cpu = {"value": t_cpu, "unit": "s", "cores": so.lib().or_get_threads(), "kind": "port",
               "sample": "full workload, one solve (oracle assembly+setup %.1fs untimed)" % t_osetup, "iterations": orr["its"]}

    cfg = workload(args, n_global)
    cfg["solver_options"] = opts
    cfg["dof_per_gpu"] = n
    line = {"metric": "time_to_solve_rtol1e-8", "value": t_solve, "unit": "s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_solve * 1e3, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": cfg, "iterations": res["its"], "converged_reason": res["reason"],
            "iterations_per_s": res["its"] / t_solve, "true_relative_residual": true_rel,
            "e2e": {"value": t_e2e, "unit": "s", "h2d_bytes_per_step": 8 * n_global, "d2h_bytes_per_step": 8 * n_global},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "kernel_classes_ms_per_solve": classes, "kernel_classes_total_ms": round(prof_total_ms, 3), "assembly_s": t_assembly, "ksp_setup_s": t_setup}
    emit(line)


if __name__ == "__main__":
    main()
