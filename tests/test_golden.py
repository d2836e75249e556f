"""Golden vectors produced by the REAL reference code (tests/golden/make_golden.py: the unmodified
/root/reference sources compiled against our PETSc shim, oracle/ref_build.sh).  They pin
  - the CPU oracle (bit-exact, CPU test),
  - the CUDA assembly path (bit-exact, GPU test),
and, where oracle/_ref exists (it is built here and travels to the GPU box), the reference is re-run live."""
import os
import subprocess
import sys

import numpy as np
import pytest

import sp_oracle as so

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
REFDIR = os.path.join(ROOT, "oracle", "_ref")
CASES = [(3, 3, "intended"), (7, 5, "intended"), (16, 16, "intended"), (3, 3, "as_written")]
FREE = [10, 11, 12, 13, 18, 19, 20, 21]
U_FREE = [0.0496585971446208, 0.0918684047175528, 0.0397268777157024, 0.0869025450030936,
          0.0397268777157024, 0.0869025450030936, 0.0496585971446208, 0.0918684047175528]


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def test_oracle_element_kernels_match_reference_bit_for_bit():
    g = np.load(os.path.join(GOLD, "ref_elements.npz"))
    assert len(g["ec"]) >= 16
    for k in range(len(g["ec"])):
        assert np.array_equal(bits(so.element_stress(g["ec"][k])), bits(g["Ke"][k])), k
        assert np.array_equal(bits(so.element_rhs(g["ec"][k], 0)), bits(g["Fe"][k])), k
    # quadrature literal (truncated on purpose, Discretization.c:52-55), weights, node order of DMDAGetElementEqnums
    assert np.array_equal(np.abs(g["gauss_xi"]), np.full((4, 2), 0.57735026919)) and np.all(g["gauss_w"] == 1.0)
    assert g["eqnums_5_7"][:, 1:].tolist() == [[7, 5, 0], [7, 5, 1], [8, 5, 0], [8, 5, 1], [8, 6, 0], [8, 6, 1], [7, 6, 0], [7, 6, 1]]
    assert np.allclose(g["Ni"].sum(axis=1), 1.0)


@pytest.mark.parametrize("nx,ny,mode", CASES)
def test_oracle_assembly_matches_reference_bit_for_bit(nx, ny, mode):
    a = np.load(os.path.join(GOLD, "ref_assembly_%dx%d_%s.npz" % (nx, ny, mode)))
    for bc in (False, True):
        t = "bc" if bc else "nobc"
        p = so.Problem(nx, ny, as_written=(mode == "as_written"), bc=bc)
        assert np.array_equal(a[t + "_rowptr"], p.A.rowptr) and np.array_equal(a[t + "_col"], p.A.col)
        assert np.array_equal(bits(a[t + "_val"]), bits(p.A.val))        # NaN payloads included (as_written)
        assert np.array_equal(bits(a[t + "_f"]), bits(p.f))


needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REFDIR, "libref_cpu.so")), reason="oracle/_ref not built (needs /root/reference)")


@needs_ref
def test_golden_files_are_what_the_reference_produces_now(tmp_path):
    sys.path.insert(0, GOLD)
    import make_golden
    names = make_golden.generate(str(tmp_path))
    assert len(names) == 5
    for n in names:
        a, b = np.load(os.path.join(GOLD, n)), np.load(os.path.join(str(tmp_path), n))
        assert sorted(a.files) == sorted(b.files)
        for k in a.files:
            x, y = a[k], b[k]
            assert np.array_equal(bits(x), bits(y)) if x.dtype == np.float64 else np.array_equal(x, y), (n, k)


def run_ref(exe, args):
    p = subprocess.run([os.path.join(REFDIR, exe)] + args.split(), cwd=str(REFDIR), capture_output=True, text=True, timeout=300)
    return p.returncode, p.stdout, p.stderr


@needs_ref
def test_unmodified_reference_main_on_the_cpu_backend():
    """BASELINE config 0: the reference's own main.c (3x3 grid) through the shim.  As written the matrix is NaN
    (the reference's GetElementCoords defect) -> KSP_DIVERGED_NANORINF; with the intended coordinates interposed it
    reproduces the known solution."""
    rc, out, err = run_ref("saddle_point_run_cpu", "-ksp_type gmres -pc_type jacobi -ksp_converged_reason")
    assert rc == 0 and "did not converge due to DIVERGED_NANORINF" in out, (out, err)
    rc, out, err = run_ref("saddle_point_run_cpu_intended", "-ksp_type gmres -pc_type jacobi -ksp_rtol 1e-12 -ksp_converged_reason -solution_view")
    assert rc == 0 and "Linear solve converged due to CONVERGED_RTOL iterations" in out, (out, err)
    vals = [float(x) for x in out.split("type: b200sp-shim")[1].split()[:32]]
    assert np.allclose(np.array(vals)[FREE], U_FREE, rtol=1e-9)
    assert os.path.exists(os.path.join(REFDIR, "test.vtk"))            # Visulaization.c ran unchanged
    head = open(os.path.join(REFDIR, "test.vtk")).read().split("\n")
    assert head[0].startswith("# vtk DataFile") and "POINTS 16 double" in head[4]


@needs_ref
def test_shim_prints_petsc_monitor_and_reason_lines():
    """-ksp_monitor / -ksp_converged_reason in PETSc's own line formats ("%3D KSP Residual norm %14.12e", the
    KSPConvergedReasons spelling); GMRES repeats the iteration number at every restart (KSPGMRESCycle monitors the
    recomputed residual at the start of a cycle)."""
    import re
    rc, out, err = run_ref("saddle_point_run_cpu_intended", "-da_grid_x 13 -da_grid_y 11 -ksp_type gmres -ksp_gmres_restart 5 -pc_type none "
                           "-ksp_rtol 1e-8 -ksp_monitor -ksp_converged_reason")
    assert rc == 0, (out, err)
    mon = re.findall(r"^\s*(\d+) KSP Residual norm (\d\.\d{12}e[-+]\d{2}) $", out, flags=re.M)
    m = re.search(r"^Linear solve converged due to CONVERGED_RTOL iterations (\d+)$", out, flags=re.M)
    assert m and mon, out
    its = int(m.group(1))
    nums = [int(a) for a, _ in mon]
    want = []
    for i in range(its + 1):
        want.append(i)
        if i % 5 == 0 and 0 < i < its:
            want.append(i)                                   # cycle start: same iteration number again
    assert nums == want, (nums, want)
    norms = [float(b) for _, b in mon]
    assert norms[-1] <= 1e-8 * norms[0] < norms[-2]
    assert all(b <= a * (1 + 1e-6) for a, b in zip(norms, norms[1:]))      # GMRES residuals never increase (recomputed ones agree to rounding)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("nx,ny,mode", CASES)
def test_cuda_assembly_matches_reference_bit_for_bit(ctx, nx, ny, mode):
    import saddle_point_petsc_b200 as sp
    a = np.load(os.path.join(GOLD, "ref_assembly_%dx%d_%s.npz" % (nx, ny, mode)))
    aw = mode == "as_written"
    da = sp.DMDA(ctx, nx, ny)
    A = da.assemble_stress(aw)
    f = sp.Vec(ctx, 2 * da.n_nodes_local)
    da.assemble_rhs(f, 0, aw)
    for t in ("nobc", "bc"):
        rp, col, val = A.csr()
        assert np.array_equal(rp, a[t + "_rowptr"]) and np.array_equal(col, a[t + "_col"])
        if aw:   # NaN positions must agree; NaN payload bits are not a meaningful target
            assert np.array_equal(np.isnan(val), np.isnan(a[t + "_val"]))
            m = ~np.isnan(val)
            assert np.array_equal(bits(val[m]), bits(a[t + "_val"][m]))
        else:
            assert np.array_equal(bits(val), bits(a[t + "_val"]))
        assert np.array_equal(bits(f.numpy()), bits(a[t + "_f"]))
        if t == "nobc":
            ids = da.bc_ids(2)
            f.set_values(ids, np.zeros(len(ids)))
            A.zero_rows_columns(ids, 1.0)


@pytest.mark.gpu
@needs_ref
def test_unmodified_reference_main_drives_the_cuda_backend():
    """The drop-in: reference main.c / SaddlePointProblem.c / Discretization.c / Visulaization.c, unmodified, linked
    against the shim over libb200sp: MatSetValuesStencil triplets -> device radix sort -> CSR, MatZeroRowsColumns
    and KSPSolve on the GPU."""
    for opts in ("-ksp_type gmres -pc_type jacobi", "-ksp_type fgmres -pc_type none", "-ksp_type minres -pc_type jacobi"):
        rc, out, err = run_ref("saddle_point_run_b200_intended", opts + " -ksp_rtol 1e-12 -ksp_converged_reason -solution_view")
        assert rc == 0 and "Linear solve converged due to CONVERGED_RTOL iterations" in out, (opts, out, err)
        vals = [float(x) for x in out.split("type: b200sp-shim")[1].split()[:32]]
        assert np.allclose(np.array(vals)[FREE], U_FREE, rtol=1e-9), opts
    # BASELINE config 0 as stated: the default main.c case with a fieldsplit-Schur preconditioned Krylov solve
    rc, out, err = run_ref("saddle_point_run_b200_intended", "-ksp_type gmres -ksp_rtol 1e-12 -pc_type fieldsplit -pc_fieldsplit_type schur "
                           "-pc_fieldsplit_schur_fact_type full -fieldsplit_0_ksp_type preonly -fieldsplit_0_pc_type jacobi "
                           "-fieldsplit_1_ksp_type preonly -fieldsplit_1_pc_type jacobi -ksp_converged_reason -solution_view")
    assert rc == 0 and "Linear solve converged due to CONVERGED_RTOL iterations" in out, (out, err)
    vals = [float(x) for x in out.split("type: b200sp-shim")[1].split()[:32]]
    assert np.allclose(np.array(vals)[FREE], U_FREE, rtol=1e-9)
    rc, out, err = run_ref("saddle_point_run_b200", "-ksp_type gmres -pc_type jacobi -ksp_converged_reason")
    assert rc == 0 and "did not converge due to DIVERGED_NANORINF" in out, (out, err)                  # as written: NaN operator, same verdict as on the CPU
    rc, out, err = run_ref("saddle_point_run_b200_intended", "-da_grid_x 33 -da_grid_y 33 -ksp_type fgmres -pc_type mg -pc_mg_levels 3 -ksp_rtol 1e-9 -ksp_converged_reason")
    assert rc == 0 and "Linear solve converged due to CONVERGED_RTOL iterations" in out, (out, err)
