"""Run only the A-block SpMV (for ncu captures): python tools/spmv_only.py [nx] [reps] [block]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import saddle_point_petsc_b200 as sp  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 2304
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
block = sys.argv[3] if len(sys.argv) > 3 else "A"
kernel = int(sys.argv[4]) if len(sys.argv) > 4 else -1
ctx = sp.Context()
da = sp.DMDA(ctx, nx, nx)
if block == "A":
    m = da.assemble_stress()
elif block in ("P", "R"):
    import ctypes as C
    h = sp._vp()
    Mc = nx // 2 + 1
    sp._chk(sp.lib().b200sp_interp_q1(ctx.h, Mc, Mc, 2, 1, C.byref(h)))
    m = sp.Mat(ctx, h)
    if block == "R":
        m = m.transpose()
else:
    Bt, B, C, Q = da.assemble_kkt()
    m = {"Bt": Bt, "B": B, "C": C}[block]
r, c, nnz = m.size()
if kernel >= 0:
    m.set_spmv_kernel(kernel)
x, y = sp.Vec(ctx, c), sp.Vec(ctx, r)
x.set(1.0)
for _ in range(3):
    m.mult(x, y)
ctx.timer_start()
for _ in range(reps):
    m.mult(x, y)
ms = ctx.timer_stop() / reps
byts = 12 * nnz + 4 * (r + 1) + 8 * r + 8 * c
fmt = m.spmv_format()
stream = fmt["matrix_bytes"] + 8 * r + 8 * c
print("block %s nx %d: %.4f ms  %.1f GB/s (algorithmic CSR bytes)  %.1f GB/s (bytes of the stored format: block %s, value_dict %s, %.2f B/nnz)"
      % (block, nx, ms, byts / ms / 1e6, stream / ms / 1e6, fmt["block"], fmt["value_dict"], fmt["matrix_bytes"] / nnz))
