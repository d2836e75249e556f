"""Generate golden vectors from the REAL reference code (oracle/_ref/libref_cpu.so = the unmodified
/root/reference sources compiled by oracle/ref_build.sh against our PETSc shim).

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

One subprocess per mode, because "intended" mode interposes our GetElementCoords (oracle/intended_coords.c) in
front of the reference object, which is process-global.  The committed .npz files are what the oracle (CPU
tests) and the CUDA path (GPU tests) are pinned against; tests/test_oracle_vs_ref.py additionally re-runs the
reference live whenever oracle/_ref exists.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFDIR = os.path.join(ROOT, "oracle", "_ref")

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


def load(intended):
    if intended:
        C.CDLL(os.path.join(REFDIR, "libintended_coords.so"), mode=C.RTLD_GLOBAL)
    return C.CDLL(os.path.join(REFDIR, "libref_cpu.so"), mode=C.RTLD_GLOBAL)


def element_vectors(L, rng):
    """FormStressOperatorQ12D / FormLaplaceRHSQ12D of the reference on squares, rectangles and random quads."""
    ecs = [np.array([0, 0, 0, 1 / 3, 1 / 3, 1 / 3, 1 / 3, 0.0]),                       # the 3x3 grid's first element
           np.array([1 / 3, 1 / 3, 1 / 3, 2 / 3, 2 / 3, 2 / 3, 2 / 3, 1 / 3]),
           np.array([0, 0, 0, 0.25, 0.5, 0.25, 0.5, 0.0])]                             # rectangle
    for _ in range(13):                                                                 # perturbed quads
        s = rng.uniform(0.01, 2.0)
        base = np.array([0, 0, 0, 1, 1, 1, 1, 0.0]) * s + rng.uniform(-1, 1)
        ecs.append(base + rng.uniform(-0.1, 0.1, 8) * s)
    ecs = np.array(ecs)
    Ke, Fe = np.zeros((len(ecs), 64)), np.zeros((len(ecs), 8))
    form_rhs = C.cast(L.FormRHS, C.c_void_p)
    L.FormStressOperatorQ12D.argtypes = [c_dp, c_dp, c_dp]
    L.FormLaplaceRHSQ12D.argtypes = [c_dp, C.c_void_p, c_dp]
    coeff = np.ones(4)
    for k in range(len(ecs)):
        ec = np.ascontiguousarray(ecs[k])
        L.FormStressOperatorQ12D(ec.ctypes.data_as(c_dp), coeff.ctypes.data_as(c_dp), Ke[k].ctypes.data_as(c_dp))
        L.FormLaplaceRHSQ12D(ec.ctypes.data_as(c_dp), form_rhs, Fe[k].ctypes.data_as(c_dp))
    # quadrature / shape functions / equation numbering, raw
    ngp = C.c_int()
    xi, w = np.zeros((4, 2)), np.zeros(4)
    L.ConstructGaussQuadratureQ12D(C.byref(ngp), xi.ctypes.data_as(c_dp), w.ctypes.data_as(c_dp))
    Ni = np.zeros((4, 4))
    for p in range(4):
        L.ConstructQ12D_Ni(xi[p].ctypes.data_as(c_dp), Ni[p].ctypes.data_as(c_dp))
    eq = np.zeros(8 * 4, dtype=np.int32)                                                 # MatStencil {k,j,i,c}
    L.DMDAGetElementEqnums(5, 7, eq.ctypes.data_as(c_ip))
    return {"ec": ecs, "Ke": Ke, "Fe": Fe, "gauss_xi": xi, "gauss_w": w, "Ni": Ni, "eqnums_5_7": eq.reshape(8, 4)}


def get_csr(L, A):
    nr, nnz = C.c_int(), C.c_int()
    assert L.MatShimGetCSR(A, C.byref(nr), C.byref(nnz), None, None, None) == 0
    rp, cj, va = np.zeros(nr.value + 1, np.int32), np.zeros(nnz.value, np.int32), np.zeros(nnz.value)
    assert L.MatShimGetCSR(A, C.byref(nr), C.byref(nnz), rp.ctypes.data_as(c_ip), cj.ctypes.data_as(c_ip), va.ctypes.data_as(c_dp)) == 0
    return rp, cj, va


def get_vec(L, f):
    n = C.c_int()
    L.VecGetSize(f, C.byref(n))
    pa = c_dp()
    L.VecGetArray(f, C.byref(pa))
    return np.ctypeslib.as_array(pa, shape=(n.value,)).copy()


def assemble(L, nx, ny):
    """SetupDMDA -> DMCreateMatrix -> AssembleOperator_Laplace -> AssembleRHS_Laplace -> ApplyBC_Laplace, exactly the
    call sequence of SolveConstraintLaplaceProblem (src/SaddlePointProblem.c:42-56), returning the CSR and f."""
    vp = C.c_void_p
    da, A, f = vp(), vp(), vp()
    assert L.PetscInitialize(None, None, None, None) == 0
    assert L.SetupDMDA(nx, ny, C.byref(da)) == 0
    assert L.DMCreateMatrix(da, C.byref(A)) == 0
    assert L.DMCreateGlobalVector(da, C.byref(f)) == 0
    assert L.AssembleOperator_Laplace(da, C.byref(A)) == 0
    assert L.AssembleRHS_Laplace(da, C.byref(f)) == 0
    out = {}
    for tag in ("nobc", "bc"):
        rp, cj, va = get_csr(L, A)
        out.update({tag + "_rowptr": rp, tag + "_col": cj, tag + "_val": va, tag + "_f": get_vec(L, f)})
        if tag == "nobc":
            assert L.ApplyBC_Laplace(da, C.byref(A), C.byref(f)) == 0
    return out


def worker(mode, outdir):
    intended = mode == "intended"
    L = load(intended)
    rng = np.random.default_rng(2026)
    if intended:
        np.savez(os.path.join(outdir, "ref_elements.npz"), **element_vectors(L, rng))
    grids = [(3, 3), (7, 5), (16, 16)] if intended else [(3, 3)]
    for nx, ny in grids:
        np.savez(os.path.join(outdir, "ref_assembly_%dx%d_%s.npz" % (nx, ny, mode)), **assemble(L, nx, ny))


def generate(outdir):
    assert os.path.exists(os.path.join(REFDIR, "libref_cpu.so")), "run oracle/ref_build.sh first (needs /root/reference)"
    for mode in ("intended", "as_written"):
        subprocess.check_call([sys.executable, os.path.abspath(__file__), mode, outdir])
    return sorted(f for f in os.listdir(outdir) if f.endswith(".npz"))


if __name__ == "__main__":
    if len(sys.argv) > 2:
        worker(sys.argv[1], sys.argv[2])
    else:
        print("golden vectors written:", generate(HERE))
