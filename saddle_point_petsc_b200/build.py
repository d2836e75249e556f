"""Build libb200sp.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m saddle_point_petsc_b200.build [--force] [--verbose]

-fmad=false: no implicit FMA contraction anywhere, so the element kernels and elementwise vector kernels
round exactly like the reference's C compiled for baseline x86-64; kernels that want an FMA say fma().
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libb200sp.so")
SOURCES = ["kernels_vec.cu", "kernels_spmv.cu", "kernels_spmv_tma.cu", "kernels_setup.cu", "kernels_assembly.cu", "kernels_assembly3d.cu", "kernels_amg.cu", "dist.cu", "dist_spgemm.cu", "solver.cu", "capi.cu"]
HEADERS = ["core.h", "dev.cuh", "solver.h", "nccl_dyn.h", "dist.h", os.path.join("..", "..", "include", "b200sp.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
         "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "-Xptxas", "-v"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    objs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    procs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(HERE, "build", s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs + [os.path.abspath(__file__)]):
            cmd = [NVCC] + FLAGS + ["-c", src, "-o", obj]
            procs.append((s, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, cmd, p in procs:
        out, _ = p.communicate()
        log = os.path.join(HERE, "build", s + ".ptxas.log")
        with open(log, "w") as f:
            f.write(out)
        if p.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed for %s:\n%s\n" % (s, out))
        elif verbose:
            print(out)
    if failed:
        raise RuntimeError("libb200sp build failed")
    if force or procs or _stale(OUT, objs):
        cmd = [NVCC, "-shared", "-Wno-deprecated-gpu-targets", "-o", OUT] + objs + ["-ldl"]
        subprocess.check_call(cmd)
    # PETSc API shim over the C ABI (plain C, host only): lets the unmodified reference sources link against us
    shim_out = os.path.join(HERE, "libpetscshim.so")
    shim_src = [os.path.join(CSRC, "petsc_shim.c"), os.path.join(CSRC, "shim_backend_b200sp.c")]
    shim_dep = shim_src + [os.path.join(CSRC, "shim_backend.h"), os.path.join(HERE, "..", "include", "petsc_shim", "petsc.h"), OUT]
    if force or _stale(shim_out, shim_dep):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-I", os.path.join(HERE, "..", "include", "petsc_shim"), "-o", shim_out]
                              + shim_src + ["-L", HERE, "-lb200sp", "-Wl,-rpath,$ORIGIN", "-lm"])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
