"""CPU test of bench.py's reference arm (the CPU oracle timed through the same harness): one JSON line on stdout with the
keys the driver reads, alone and under torchrun with two ranks (rank 0 prints, the other rank exits 0 silently)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "cpu_baseline", "e2e"}


def check(line, n):
    d = json.loads(line)
    assert KEYS <= set(d), KEYS - set(d)
    assert d["impl"] == "reference" and d["n_gpus"] == n and d["steps"] == 1 and d["higher_is_better"] is False
    assert d["metric"] == "time_to_solve_rtol1e-8" and d["unit"] == "s" and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("kkt2d_") and d["config"]["grid_elements"] == [32, 32]
    assert d["value"] > 0 and abs(d["ms_per_step"] - 1e3 * d["value"]) < 1e-6 * d["ms_per_step"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["converged_reason"] == 2 and 10 <= d["iterations"] <= 20


def test_reference_arm_prints_one_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--nx", "32", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    check(lines[0], 1)


def test_reference_arm_under_torchrun_two_ranks():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29655", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--nx", "32", "--steps", "1",
           "--warmup", "0"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1, p.stdout
    check(lines[0], 2)
