"""saddle_point_petsc_b200 -- B200-native (sm_100a) Schur/Krylov hot path of p-m-mueller/saddle_point_petsc.

Python host-side mirror of the PETSc objects the reference drives (DM / Mat / Vec / KSP;
src/SaddlePointProblem.c:34-76), as thin ctypes wrappers over the C ABI in include/b200sp.h.
There is no CPU fallback: if libb200sp.so is missing this module raises on first use, and every
compute call fails with B200SP_ERR_NO_DEVICE when no B200 is visible.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200sp.so")
_lib = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
_vp = C.c_void_p

ERR_NAMES = {1: "ARG", 2: "NO_DEVICE", 3: "CUDA", 4: "UNSUPPORTED", 5: "NCCL", 6: "MEM"}
REASONS = {2: "CONVERGED_RTOL", 3: "CONVERGED_ATOL", 4: "CONVERGED_ITS", -3: "DIVERGED_ITS", -4: "DIVERGED_DTOL",
           -5: "DIVERGED_BREAKDOWN", -8: "DIVERGED_INDEFINITE_PC", -9: "DIVERGED_NANORINF"}


class B200spError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("b200sp error %d (%s): %s" % (code, ERR_NAMES.get(code, "?"), msg))
        self.code = code


# (name, argtypes) -- every symbol include/b200sp.h declares; restype is int unless listed in _NONINT
_SIGS = {
    "b200sp_nccl_unique_id": [C.c_char_p],
    "b200sp_ctx_create": [C.c_int, C.c_int, C.c_int, C.c_char_p, C.POINTER(_vp)],
    "b200sp_local_group_create": [C.c_int, C.POINTER(_vp)],
    "b200sp_local_group_destroy": [_vp],
    "b200sp_ctx_create_local": [_vp, C.c_int, C.c_int, C.POINTER(_vp)],
    "b200sp_ctx_destroy": [_vp],
    "b200sp_ctx_synchronize": [_vp],
    "b200sp_ctx_get_stream": [_vp, C.POINTER(_vp)],
    "b200sp_ctx_get_launch_count": [_vp, C.POINTER(C.c_int64)],
    "b200sp_ctx_timer_start": [_vp],
    "b200sp_ctx_timer_stop": [_vp, c_dp],
    "b200sp_ctx_profile_enable": [_vp, C.c_int],
    "b200sp_ctx_profile_report": [_vp, C.c_char_p, C.c_int],
    "b200sp_dmda_proc_grid": [C.c_int, C.c_int, C.c_int, c_ip, c_ip],
    "b200sp_dmda_ownership": [C.c_int, C.c_int, c_ip],
    "b200sp_dmda_corners": [C.c_int, C.c_int, C.c_int, C.c_int, c_ip, c_ip, c_ip, c_ip],
    "b200sp_dmda_element_corners": [C.c_int, C.c_int, C.c_int, C.c_int, c_ip, c_ip, c_ip, c_ip],
    "b200sp_dmda_global_node": [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_ip, c_ip],
    "b200sp_dmda_halo_plan": [C.c_int, C.c_int, C.c_int, C.c_int, c_ip, c_ip, c_ip, c_ip, c_ip, c_ip],
    "b200sp_dmda_halo_push_table": [C.c_int, C.c_int, C.c_int, C.c_int, c_ip, c_ip, c_ip, c_ip, c_ip],
    "b200sp_dmda_create": [_vp, C.c_int, C.c_int, C.POINTER(_vp)],
    "b200sp_dmda_destroy": [_vp],
    "b200sp_dmda_get_info": [_vp, c_ip, c_ip, c_ip, c_ip, c_ip, c_ip],
    "b200sp_dmda_bc_ids": [_vp, C.c_int, c_ip, c_ip],
    "b200sp_vec_create": [_vp, C.c_int64, C.POINTER(_vp)],
    "b200sp_vec_destroy": [_vp],
    "b200sp_vec_get_size": [_vp, C.POINTER(C.c_int64)],
    "b200sp_vec_set": [_vp, C.c_double],
    "b200sp_vec_set_values_host": [_vp, C.c_int64, c_ip, c_dp],
    "b200sp_vec_copy_from_host": [_vp, c_dp, C.c_int64],
    "b200sp_vec_copy_to_host": [_vp, c_dp, C.c_int64],
    "b200sp_vec_get_device_ptr": [_vp, C.POINTER(c_dp)],
    "b200sp_vec_copy": [_vp, _vp],
    "b200sp_vec_scale": [_vp, C.c_double],
    "b200sp_vec_axpy": [_vp, C.c_double, _vp],
    "b200sp_vec_aypx": [_vp, C.c_double, _vp],
    "b200sp_vec_waxpy": [_vp, C.c_double, _vp, _vp],
    "b200sp_vec_pointwise_mult": [_vp, _vp, _vp],
    "b200sp_vec_dot": [_vp, _vp, c_dp],
    "b200sp_vec_norm": [_vp, c_dp],
    "b200sp_vec_mdot": [_vp, C.c_int, C.POINTER(_vp), c_dp],
    "b200sp_vec_maxpy": [_vp, C.c_int, c_dp, C.POINTER(_vp)],
    "b200sp_bench_orthogonalization": [_vp, C.c_int64, C.c_int, C.c_int, c_dp, c_dp],
    "b200sp_mat_create_csr": [_vp, C.c_int, C.c_int, c_ip, c_ip, c_dp, C.POINTER(_vp)],
    "b200sp_mat_create_coo": [_vp, C.c_int, C.c_int, C.c_int64, c_ip, c_ip, c_dp, C.POINTER(_vp)],
    "b200sp_mat_destroy": [_vp],
    "b200sp_mat_set_grid": [_vp, C.c_int, C.c_int, C.c_int],
    "b200sp_mat_get_size": [_vp, c_ip, c_ip, C.POINTER(C.c_int64)],
    "b200sp_mat_get_csr_host": [_vp, c_ip, c_ip, c_dp],
    "b200sp_mat_get_spmv_plan": [_vp, C.POINTER(C.c_int64), c_ip, c_ip],
    "b200sp_mat_get_spmv_format": [_vp, c_ip, c_ip, c_ip, C.POINTER(C.c_int64)],
    "b200sp_mat_set_spmv_kernel": [_vp, C.c_int],
    "b200sp_mat_set_spmv_format": [_vp, C.c_int, C.c_int],
    "b200sp_mat_mult": [_vp, _vp, _vp],
    "b200sp_mat_mult_add": [_vp, _vp, _vp, _vp],
    "b200sp_mat_residual": [_vp, _vp, _vp, _vp],
    "b200sp_mat_get_diagonal": [_vp, _vp],
    "b200sp_mat_transpose": [_vp, C.POINTER(_vp)],
    "b200sp_mat_matmult": [_vp, _vp, C.POINTER(_vp)],
    "b200sp_mat_scale_columns": [_vp, _vp, C.POINTER(_vp)],
    "b200sp_mat_add_scaled": [_vp, C.c_double, _vp, C.POINTER(_vp)],
    "b200sp_amg_aggregate": [_vp, C.c_int, C.c_double, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)],
    "b200sp_amg_prolongator": [_vp, C.c_int, C.c_double, C.c_int, C.c_double, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(_vp)],
    "b200sp_mat_zero_rows_columns": [_vp, C.c_int, c_ip, C.c_double],
    "b200sp_mat_zero_rows": [_vp, C.c_int, c_ip, C.c_double],
    "b200sp_mat_zero_columns": [_vp, C.c_int, c_ip],
    "b200sp_mat_create_nest": [_vp, _vp, _vp, _vp, C.POINTER(_vp)],
    "b200sp_assemble_stress": [_vp, C.c_int, C.POINTER(_vp)],
    "b200sp_assemble_stress_coeff": [_vp, C.c_int, C.POINTER(_vp)],
    "b200sp_assemble_rhs": [_vp, C.c_int, C.c_int, _vp],
    "b200sp_assemble_kkt": [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)],
    "b200sp_dmda3d_proc_grid": [C.c_int, C.c_int, C.c_int, C.c_int, c_ip, c_ip, c_ip],
    "b200sp_dmda3d_corners": [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_ip, c_ip, c_ip, c_ip, c_ip, c_ip],
    "b200sp_dmda3d_global_node": [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_ip, c_ip],
    "b200sp_dmda3d_create": [_vp, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)],
    "b200sp_dmda3d_destroy": [_vp],
    "b200sp_dmda3d_get_info": [_vp, c_ip, c_ip, c_ip, c_ip, c_ip, c_ip, C.POINTER(C.c_int64)],
    "b200sp_dmda3d_bc_ids": [_vp, C.c_int, c_ip, c_ip],
    "b200sp_assemble3d_stress": [_vp, C.POINTER(_vp)],
    "b200sp_assemble3d_rhs": [_vp, C.c_int, _vp],
    "b200sp_assemble3d_kkt": [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)],
    "b200sp_assemble_constraints": [_vp, C.POINTER(_vp), C.POINTER(_vp)],
    "b200sp_interp_q1": [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)],
    "b200sp_mat_mult_transpose": [_vp, _vp, _vp],
    "b200sp_pc_create": [_vp, C.POINTER(_vp)],
    "b200sp_pc_destroy": [C.POINTER(_vp)],
    "b200sp_pc_set_operators": [_vp, _vp, _vp],
    "b200sp_pc_set_options": [_vp, C.c_char_p],
    "b200sp_pc_set_schur_user_mat": [_vp, _vp],
    "b200sp_pc_set_dmda": [_vp, _vp],
    "b200sp_pc_setup": [_vp],
    "b200sp_pc_apply": [_vp, _vp, _vp],
    "b200sp_pc_view": [_vp, C.c_char_p, C.c_int],
    "b200sp_ksp_create": [_vp, C.POINTER(_vp)],
    "b200sp_ksp_destroy": [C.POINTER(_vp)],
    "b200sp_ksp_set_operators": [_vp, _vp, _vp],
    "b200sp_ksp_set_options": [_vp, C.c_char_p],
    "b200sp_ksp_set_schur_user_mat": [_vp, _vp],
    "b200sp_ksp_set_dmda": [_vp, _vp],
    "b200sp_ksp_setup": [_vp],
    "b200sp_ksp_solve": [_vp, _vp, _vp],
    "b200sp_ksp_solve_host": [_vp, c_dp, c_dp, C.c_int64],
    "b200sp_ksp_get_iteration_number": [_vp, c_ip],
    "b200sp_ksp_get_residual_norm": [_vp, c_dp],
    "b200sp_ksp_get_converged_reason": [_vp, c_ip],
    "b200sp_ksp_get_residual_history": [_vp, c_dp, C.c_int, c_ip],
    "b200sp_ksp_pc_apply": [_vp, _vp, _vp],
    "b200sp_ksp_view": [_vp, C.c_char_p, C.c_int],
}
_NONINT = {"b200sp_last_error": C.c_char_p, "b200sp_version": C.c_char_p}
ABI_SYMBOLS = sorted(list(_SIGS) + list(_NONINT))


def lib():
    """Load libb200sp.so (built by saddle_point_petsc_b200.build).  Raises if it is missing: no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libb200sp.so not built: run `python -m saddle_point_petsc_b200.build` "
                              "(there is no CPU fallback for the solver)")
        L = C.CDLL(LIB_PATH)
        for name, args in _SIGS.items():
            f = getattr(L, name)
            f.restype = C.c_int
            f.argtypes = args
        for name, res in _NONINT.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = []
        _lib = L
    return _lib


def _chk(code):
    if code != 0:
        raise B200spError(code, lib().b200sp_last_error().decode())


def _dptr(a):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(c_dp)


def _iptr(a):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(c_ip)


# ---------------------------------------------------------------- host-only DMDA helpers (work without a GPU)
def dmda_proc_grid(M, N, size):
    m, n = C.c_int(), C.c_int()
    _chk(lib().b200sp_dmda_proc_grid(M, N, size, C.byref(m), C.byref(n)))
    return m.value, n.value


def dmda_ownership(M, m):
    lx = np.zeros(m, dtype=np.int32)
    _chk(lib().b200sp_dmda_ownership(M, m, _iptr(lx)))
    return lx


def dmda_corners(M, N, size, rank):
    v = [C.c_int() for _ in range(4)]
    _chk(lib().b200sp_dmda_corners(M, N, size, rank, *[C.byref(x) for x in v]))
    return tuple(x.value for x in v)


def dmda_element_corners(M, N, size, rank):
    v = [C.c_int() for _ in range(4)]
    _chk(lib().b200sp_dmda_element_corners(M, N, size, rank, *[C.byref(x) for x in v]))
    return tuple(x.value for x in v)


def dmda_global_node(M, N, size, i, j):
    g, o = C.c_int(), C.c_int()
    _chk(lib().b200sp_dmda_global_node(M, N, size, i, j, C.byref(g), C.byref(o)))
    return g.value, o.value


def dmda_halo_plan(M, N, size, rank):
    ng, ns = C.c_int(), C.c_int()
    _chk(lib().b200sp_dmda_halo_plan(M, N, size, rank, C.byref(ng), None, None, C.byref(ns), None, None))
    gg, go = np.zeros(ng.value, dtype=np.int32), np.zeros(ng.value, dtype=np.int32)
    sr, sl = np.zeros(ns.value, dtype=np.int32), np.zeros(ns.value, dtype=np.int32)
    _chk(lib().b200sp_dmda_halo_plan(M, N, size, rank, C.byref(ng), _iptr(gg), _iptr(go), C.byref(ns), _iptr(sr), _iptr(sl)))
    return {"ghost_gnode": gg, "ghost_owner": go, "send_rank": sr, "send_lnode": sl}


def dmda_halo_push_table(M, N, size, rank):
    """Node-keyed send table (fused halo push): node_ent per owned node, entry -> (destination rank, position)."""
    no, ne = C.c_int(), C.c_int()
    _chk(lib().b200sp_dmda_halo_push_table(M, N, size, rank, C.byref(no), None, C.byref(ne), None, None))
    ent = np.zeros(no.value, dtype=np.int32)
    er, ep = np.zeros(ne.value, dtype=np.int32), np.zeros(ne.value, dtype=np.int32)
    _chk(lib().b200sp_dmda_halo_push_table(M, N, size, rank, C.byref(no), _iptr(ent), C.byref(ne), _iptr(er), _iptr(ep)))
    return {"node_ent": ent, "entry_rank": er, "entry_pos": ep}


# ---------------------------------------------------------------- objects
class Context:
    """PetscInitialize + PETSC_COMM_WORLD: one per process per GPU."""

    def __init__(self, device=0, rank=0, size=1, nccl_id=None):
        self.h = _vp()
        _chk(lib().b200sp_ctx_create(device, rank, size, nccl_id, C.byref(self.h)))
        self.rank, self.size, self.device = rank, size, device

    @classmethod
    def local(cls, group, rank, device=0):
        """rank `rank` of an in-process LocalGroup (call from the thread that will drive this rank)"""
        self = cls.__new__(cls)
        self.h = _vp()
        _chk(lib().b200sp_ctx_create_local(group.h, rank, device, C.byref(self.h)))
        self.rank, self.size, self.device = rank, group.size, device
        return self

    @staticmethod
    def nccl_unique_id():
        buf = C.create_string_buffer(128)
        _chk(lib().b200sp_nccl_unique_id(buf))
        return buf.raw

    def synchronize(self):
        _chk(lib().b200sp_ctx_synchronize(self.h))

    def launch_count(self):
        n = C.c_int64()
        _chk(lib().b200sp_ctx_get_launch_count(self.h, C.byref(n)))
        return n.value

    def timer_start(self):
        _chk(lib().b200sp_ctx_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_double()
        _chk(lib().b200sp_ctx_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def profile(self, on):
        _chk(lib().b200sp_ctx_profile_enable(self.h, int(on)))

    def profile_report(self):
        import json
        buf = C.create_string_buffer(1 << 16)
        _chk(lib().b200sp_ctx_profile_report(self.h, buf, len(buf)))
        return json.loads(buf.value.decode())

    def bench_orthogonalization(self, n, k, reps=10):
        """(ms per VecMDot launch, ms per VecMAXPY+norm launch) of the fused GMRES kernels on k basis vectors"""
        a, b = C.c_double(), C.c_double()
        _chk(lib().b200sp_bench_orthogonalization(self.h, n, k, reps, C.byref(a), C.byref(b)))
        return a.value, b.value

    def destroy(self):
        if self.h:
            _chk(lib().b200sp_ctx_destroy(self.h))
            self.h = _vp()


class LocalGroup:
    """All ranks as threads of this process (b200sp_local_group_create); see run_ranks()."""

    def __init__(self, size):
        self.size = size
        self.h = _vp()
        _chk(lib().b200sp_local_group_create(size, C.byref(self.h)))


def run_ranks(size, fn, device=0):
    """Run fn(ctx) SPMD-style on `size` rank-threads of one process (ctypes releases the GIL inside the library, and
    every collective is a host barrier, so the ranks make progress together).  Returns [fn result per rank]."""
    import threading
    grp = LocalGroup(size)
    out, err = [None] * size, [None] * size

    def work(r):
        try:
            ctx = Context.local(grp, r, device)
            out[r] = fn(ctx)
        except BaseException as e:  # noqa: BLE001 - reported to the caller below
            err[r] = e

    ts = [threading.Thread(target=work, args=(r,)) for r in range(size)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for e in err:
        if e is not None:
            raise e
    return out


class Vec:
    def __init__(self, ctx, n=None, handle=None):
        self.ctx = ctx
        if handle is None:
            handle = _vp()
            _chk(lib().b200sp_vec_create(ctx.h, n, C.byref(handle)))
        self.h = handle
        sz = C.c_int64()
        _chk(lib().b200sp_vec_get_size(self.h, C.byref(sz)))
        self.n = sz.value

    @classmethod
    def from_numpy(cls, ctx, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        v = cls(ctx, len(a))
        _chk(lib().b200sp_vec_copy_from_host(v.h, _dptr(a), len(a)))
        return v

    def numpy(self):
        out = np.empty(self.n)
        _chk(lib().b200sp_vec_copy_to_host(self.h, _dptr(out), self.n))
        return out

    def set(self, a):
        _chk(lib().b200sp_vec_set(self.h, float(a)))

    def set_values(self, idx, vals):
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        _chk(lib().b200sp_vec_set_values_host(self.h, len(idx), _iptr(idx), _dptr(vals)))

    def copy_to(self, y):
        _chk(lib().b200sp_vec_copy(self.h, y.h))

    def scale(self, a):
        _chk(lib().b200sp_vec_scale(self.h, float(a)))

    def axpy(self, a, x):
        _chk(lib().b200sp_vec_axpy(self.h, float(a), x.h))

    def aypx(self, a, x):
        _chk(lib().b200sp_vec_aypx(self.h, float(a), x.h))

    def waxpy(self, a, x, y):
        _chk(lib().b200sp_vec_waxpy(self.h, float(a), x.h, y.h))

    def pointwise_mult(self, x, y):
        _chk(lib().b200sp_vec_pointwise_mult(self.h, x.h, y.h))

    def dot(self, y):
        r = C.c_double()
        _chk(lib().b200sp_vec_dot(self.h, y.h, C.byref(r)))
        return r.value

    def norm(self):
        r = C.c_double()
        _chk(lib().b200sp_vec_norm(self.h, C.byref(r)))
        return r.value

    def mdot(self, ys):
        out = np.zeros(len(ys))
        arr = (_vp * len(ys))(*[y.h for y in ys])
        _chk(lib().b200sp_vec_mdot(self.h, len(ys), arr, _dptr(out)))
        return out

    def maxpy(self, a, xs):
        a = np.ascontiguousarray(a, dtype=np.float64)
        arr = (_vp * len(xs))(*[x.h for x in xs])
        _chk(lib().b200sp_vec_maxpy(self.h, len(xs), _dptr(a), arr))

    def destroy(self):
        if self.h:
            _chk(lib().b200sp_vec_destroy(self.h))
            self.h = _vp()


class Mat:
    def __init__(self, ctx, handle):
        self.ctx = ctx
        self.h = handle
        self.blocks = None

    @classmethod
    def from_csr(cls, ctx, nrows, ncols, rowptr, col, val):
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
        col = np.ascontiguousarray(col, dtype=np.int32)
        val = np.ascontiguousarray(val, dtype=np.float64)
        h = _vp()
        _chk(lib().b200sp_mat_create_csr(ctx.h, nrows, ncols, _iptr(rowptr), _iptr(col), _dptr(val), C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def from_scipy(cls, ctx, A):
        A = A.tocsr()
        A.sort_indices()
        return cls.from_csr(ctx, A.shape[0], A.shape[1], A.indptr, A.indices, A.data)

    @classmethod
    def from_coo(cls, ctx, nrows, ncols, row, col, val):
        row = np.ascontiguousarray(row, dtype=np.int32)
        col = np.ascontiguousarray(col, dtype=np.int32)
        val = np.ascontiguousarray(val, dtype=np.float64)
        h = _vp()
        _chk(lib().b200sp_mat_create_coo(ctx.h, nrows, ncols, len(row), _iptr(row), _iptr(col), _dptr(val), C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def nest(cls, A00, A01, A10, A11=None):
        h = _vp()
        _chk(lib().b200sp_mat_create_nest(A00.h, A01.h, A10.h, A11.h if A11 is not None else None, C.byref(h)))
        m = cls(A00.ctx, h)
        m.blocks = (A00, A01, A10, A11)
        return m

    def size(self):
        r, c, z = C.c_int(), C.c_int(), C.c_int64()
        _chk(lib().b200sp_mat_get_size(self.h, C.byref(r), C.byref(c), C.byref(z)))
        return r.value, c.value, z.value

    def csr(self):
        nrows, ncols, nnz = self.size()
        rowptr = np.zeros(nrows + 1, dtype=np.int32)
        col = np.zeros(nnz, dtype=np.int32)
        val = np.zeros(nnz)
        _chk(lib().b200sp_mat_get_csr_host(self.h, _iptr(rowptr), _iptr(col), _dptr(val)))
        return rowptr, col, val

    def scipy(self):
        import scipy.sparse as sp
        nrows, ncols, _ = self.size()
        rowptr, col, val = self.csr()
        return sp.csr_matrix((val, col, rowptr), shape=(nrows, ncols))

    def spmv_plan(self):
        hist = (C.c_int64 * 14)()
        k, mr = C.c_int(), C.c_int()
        _chk(lib().b200sp_mat_get_spmv_plan(self.h, hist, C.byref(k), C.byref(mr)))
        return {"hist": list(hist), "kernel": k.value, "max_row_nnz": mr.value}

    def spmv_format(self):
        """What the TMA SpMV streams for this matrix (valid after the first mult)."""
        br, bc, vd, nb = C.c_int(), C.c_int(), C.c_int(), C.c_int64()
        _chk(lib().b200sp_mat_get_spmv_format(self.h, C.byref(br), C.byref(bc), C.byref(vd), C.byref(nb)))
        return {"block": (br.value, bc.value), "value_dict": bool(vd.value), "matrix_bytes": nb.value}

    def set_spmv_format(self, block_index=True, value_dict=True):
        """Restrict what the SpMV derives from the CSR arrays (general-matrix paths for benchmarks; same bits)."""
        _chk(lib().b200sp_mat_set_spmv_format(self.h, int(block_index), int(value_dict)))

    def set_spmv_kernel(self, k):
        _chk(lib().b200sp_mat_set_spmv_kernel(self.h, k))

    def mult(self, x, y):
        _chk(lib().b200sp_mat_mult(self.h, x.h, y.h))

    def mult_transpose(self, x, y):
        _chk(lib().b200sp_mat_mult_transpose(self.h, x.h, y.h))

    def mult_add(self, x, y, z):
        _chk(lib().b200sp_mat_mult_add(self.h, x.h, y.h, z.h))

    def residual(self, b, x, r):
        _chk(lib().b200sp_mat_residual(self.h, b.h, x.h, r.h))

    def get_diagonal(self, d):
        _chk(lib().b200sp_mat_get_diagonal(self.h, d.h))

    def transpose(self):
        h = _vp()
        _chk(lib().b200sp_mat_transpose(self.h, C.byref(h)))
        return Mat(self.ctx, h)

    def matmult(self, B):
        h = _vp()
        _chk(lib().b200sp_mat_matmult(self.h, B.h, C.byref(h)))
        return Mat(self.ctx, h)

    def scale_columns(self, d):
        h = _vp()
        _chk(lib().b200sp_mat_scale_columns(self.h, d.h, C.byref(h)))
        return Mat(self.ctx, h)

    def add_scaled(self, s, B):
        h = _vp()
        _chk(lib().b200sp_mat_add_scaled(self.h, float(s), B.h, C.byref(h)))
        return Mat(self.ctx, h)

    def amg_aggregate(self, bs=1, theta=0.0, order="hash"):
        """(aggregate id per bs-dof node, -1 = left out; number of aggregates) of the -pc_type gamg set-up."""
        nrows, _, _ = self.size()
        agg = np.zeros(max(nrows // bs, 1), dtype=np.int32)
        nagg = C.c_int()
        _chk(lib().b200sp_amg_aggregate(self.h, int(bs), float(theta), {"hash": 0, "natural": 1}[order], _iptr(agg), C.byref(nagg)))
        return agg[:nrows // bs], nagg.value

    def amg_prolongator(self, bs=1, theta=0.0, omega=0.0, node_weight=None, nagg=None, order="hash"):
        """Tentative (omega = 0) or smoothed aggregation prolongator P_t - omega D^-1 A P_t; with nagg given also
        returns the coarse node weights (finest-level nodes per aggregate)."""
        h = _vp()
        w = None if node_weight is None else np.ascontiguousarray(node_weight, dtype=np.int32)
        wc = None if nagg is None else np.zeros(max(nagg, 1), dtype=np.int32)
        _chk(lib().b200sp_amg_prolongator(self.h, int(bs), float(theta), {"hash": 0, "natural": 1}[order], float(omega), None if w is None else _iptr(w),
                                          None if wc is None else _iptr(wc), C.byref(h)))
        return Mat(self.ctx, h) if nagg is None else (Mat(self.ctx, h), wc[:nagg])

    def zero_rows_columns(self, rows, diag=1.0):
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        _chk(lib().b200sp_mat_zero_rows_columns(self.h, len(rows), _iptr(rows), diag))

    def zero_rows(self, rows, diag=0.0):
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        _chk(lib().b200sp_mat_zero_rows(self.h, len(rows), _iptr(rows), diag))

    def zero_columns(self, cols):
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        _chk(lib().b200sp_mat_zero_columns(self.h, len(cols), _iptr(cols)))

    def destroy(self):
        if self.h:
            _chk(lib().b200sp_mat_destroy(self.h))
            self.h = _vp()


class DMDA:
    """DMDACreate2d(..., nx+1, ny+1, dof, box stencil width 1) (src/Discretization.c:17)."""

    def __init__(self, ctx, nx, ny):
        self.ctx = ctx
        self.M, self.N = nx + 1, ny + 1
        self.h = _vp()
        _chk(lib().b200sp_dmda_create(ctx.h, self.M, self.N, C.byref(self.h)))
        v = [C.c_int() for _ in range(6)]
        _chk(lib().b200sp_dmda_get_info(self.h, *[C.byref(x) for x in v]))
        _, _, self.xs, self.ys, self.xm, self.ym = [x.value for x in v]
        self.n_nodes_local = self.xm * self.ym

    def bc_ids(self, dof=2):
        n = C.c_int()
        _chk(lib().b200sp_dmda_bc_ids(self.h, dof, C.byref(n), None))
        ids = np.zeros(n.value, dtype=np.int32)
        _chk(lib().b200sp_dmda_bc_ids(self.h, dof, C.byref(n), _iptr(ids)))
        return ids

    def assemble_stress(self, as_written=False):
        h = _vp()
        _chk(lib().b200sp_assemble_stress(self.h, int(as_written), C.byref(h)))
        return Mat(self.ctx, h)

    def assemble_stress_coeff(self, coeff_kind=1):
        """A with a coefficient per Gauss point (1 = smooth viscosity 1 + x(1-y)/2)"""
        h = _vp()
        _chk(lib().b200sp_assemble_stress_coeff(self.h, coeff_kind, C.byref(h)))
        return Mat(self.ctx, h)

    def assemble_rhs(self, f, rhs_kind=0, as_written=False):
        _chk(lib().b200sp_assemble_rhs(self.h, int(as_written), rhs_kind, f.h))

    def assemble_kkt(self):
        hs = [_vp() for _ in range(4)]
        _chk(lib().b200sp_assemble_kkt(self.h, *[C.byref(h) for h in hs]))
        return tuple(Mat(self.ctx, h) for h in hs)

    def assemble_constraints(self):
        """AssembleOperator_Constraints (stub in the reference): the 4 dense constraint rows B and B^T"""
        hb, hbt = _vp(), _vp()
        _chk(lib().b200sp_assemble_constraints(self.h, C.byref(hb), C.byref(hbt)))
        return Mat(self.ctx, hb), Mat(self.ctx, hbt)

    def destroy(self):
        if self.h:
            _chk(lib().b200sp_dmda_destroy(self.h))
            self.h = _vp()


class DMDA3D:
    """DMDACreate3d(..., nx+1, ny+1, nz+1, box stencil width 1): the 3-D grid of BASELINE config 4."""

    def __init__(self, ctx, nx, ny, nz):
        self.ctx = ctx
        self.M, self.N, self.P = nx + 1, ny + 1, nz + 1
        self.h = _vp()
        _chk(lib().b200sp_dmda3d_create(ctx.h, self.M, self.N, self.P, C.byref(self.h)))
        v = [C.c_int() for _ in range(6)]
        g0 = C.c_int64()
        _chk(lib().b200sp_dmda3d_get_info(self.h, *[C.byref(x) for x in v], C.byref(g0)))
        self.xs, self.ys, self.zs, self.xm, self.ym, self.zm = [x.value for x in v]
        self.gstart = g0.value
        self.n_nodes_local = self.xm * self.ym * self.zm

    def bc_ids(self, dof=3):
        n = C.c_int()
        _chk(lib().b200sp_dmda3d_bc_ids(self.h, dof, C.byref(n), None))
        ids = np.zeros(n.value, dtype=np.int32)
        _chk(lib().b200sp_dmda3d_bc_ids(self.h, dof, C.byref(n), _iptr(ids)))
        return ids

    def assemble_stress(self):
        h = _vp()
        _chk(lib().b200sp_assemble3d_stress(self.h, C.byref(h)))
        return Mat(self.ctx, h)

    def assemble_rhs(self, f, rhs_kind=1):
        _chk(lib().b200sp_assemble3d_rhs(self.h, rhs_kind, f.h))

    def assemble_kkt(self):
        hs = [_vp() for _ in range(4)]
        _chk(lib().b200sp_assemble3d_kkt(self.h, *[C.byref(h) for h in hs]))
        return tuple(Mat(self.ctx, h) for h in hs)

    def destroy(self):
        if self.h:
            _chk(lib().b200sp_dmda3d_destroy(self.h))
            self.h = _vp()


class PC:
    """PCCreate / PCSetOperators / PCSetFromOptions / PCSetUp / PCApply as an object of its own."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.h = _vp()
        _chk(lib().b200sp_pc_create(ctx.h, C.byref(self.h)))
        self._keep = []

    def set_operators(self, A, P=None):
        P = A if P is None else P
        self._keep += [A, P]
        _chk(lib().b200sp_pc_set_operators(self.h, A.h, P.h))

    def set_options(self, text):
        _chk(lib().b200sp_pc_set_options(self.h, text.encode()))

    def set_schur_user_mat(self, Q):
        self._keep.append(Q)
        _chk(lib().b200sp_pc_set_schur_user_mat(self.h, Q.h))

    def set_dmda(self, da):
        _chk(lib().b200sp_pc_set_dmda(self.h, da.h))

    def setup(self):
        _chk(lib().b200sp_pc_setup(self.h))

    def apply(self, x, y):
        _chk(lib().b200sp_pc_apply(self.h, x.h, y.h))

    def view(self):
        buf = C.create_string_buffer(1 << 16)
        _chk(lib().b200sp_pc_view(self.h, buf, len(buf)))
        return buf.value.decode()

    def destroy(self):
        if self.h:
            _chk(lib().b200sp_pc_destroy(C.byref(self.h)))


class KSP:
    """KSPCreate / KSPSetOperators / KSPSetFromOptions / KSPSetUp / KSPSolve (src/SaddlePointProblem.c:65-72)."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.h = _vp()
        _chk(lib().b200sp_ksp_create(ctx.h, C.byref(self.h)))
        self._keep = []

    def set_operators(self, A, P=None):
        P = A if P is None else P
        self._keep += [A, P]
        _chk(lib().b200sp_ksp_set_operators(self.h, A.h, P.h))

    def set_options(self, text):
        _chk(lib().b200sp_ksp_set_options(self.h, text.encode()))

    def set_schur_user_mat(self, Q):
        self._keep.append(Q)
        _chk(lib().b200sp_ksp_set_schur_user_mat(self.h, Q.h))

    def set_dmda(self, da):
        _chk(lib().b200sp_ksp_set_dmda(self.h, da.h))

    def setup(self):
        _chk(lib().b200sp_ksp_setup(self.h))

    def solve(self, b, x):
        _chk(lib().b200sp_ksp_solve(self.h, b.h, x.h))
        return self.result()

    def solve_host(self, b, x):
        assert b.dtype == np.float64 and x.dtype == np.float64 and len(b) == len(x)
        _chk(lib().b200sp_ksp_solve_host(self.h, _dptr(b), _dptr(x), len(b)))
        return self.result()

    def pc_apply(self, x, y):
        _chk(lib().b200sp_ksp_pc_apply(self.h, x.h, y.h))

    def result(self):
        its, reason, rn, ln = C.c_int(), C.c_int(), C.c_double(), C.c_int()
        _chk(lib().b200sp_ksp_get_iteration_number(self.h, C.byref(its)))
        _chk(lib().b200sp_ksp_get_converged_reason(self.h, C.byref(reason)))
        _chk(lib().b200sp_ksp_get_residual_norm(self.h, C.byref(rn)))
        _chk(lib().b200sp_ksp_get_residual_history(self.h, None, 0, C.byref(ln)))
        hist = np.zeros(max(ln.value, 1))
        _chk(lib().b200sp_ksp_get_residual_history(self.h, _dptr(hist), ln.value, C.byref(ln)))
        return {"its": its.value, "reason": reason.value, "rnorm": rn.value, "history": hist[:ln.value]}

    def view(self):
        buf = C.create_string_buffer(1 << 16)
        _chk(lib().b200sp_ksp_view(self.h, buf, len(buf)))
        return buf.value.decode()

    def destroy(self):
        if self.h:
            _chk(lib().b200sp_ksp_destroy(C.byref(self.h)))


class SaddlePointProblem:
    """Device-side equivalent of SolveConstraintLaplaceProblem (src/SaddlePointProblem.c:34-76): DMDA ->
    assemble A, f -> Dirichlet BC -> (kkt=True: B^T, B, C, Q blocks and the 2x2 nest) -> KSP."""

    def __init__(self, ctx, nx, ny, kkt=False, rhs_kind=0, as_written=False, constraints=False, g=(0.0, 0.0, 0.0, 0.0)):
        self.ctx = ctx
        self.da = DMDA(ctx, nx, ny)
        da = self.da
        self.nu, self.np_ = 2 * da.n_nodes_local, da.n_nodes_local
        self.kkt = kkt
        self.A = da.assemble_stress(as_written)
        self.constraints = constraints
        if constraints:   # the reference's own [A Bt; B 0] with 4 constraint rows (src/SaddlePointProblem.c:45-60)
            self.np_ = 4
        n = self.nu + (self.np_ if (kkt or constraints) else 0)
        self.rhs = Vec(ctx, n)
        da.assemble_rhs(self.rhs, rhs_kind, as_written)   # fills the velocity part, pressure part stays 0 (g = 0)
        self.bc = da.bc_ids(2)
        self.rhs.set_values(self.bc, np.zeros(len(self.bc)))
        self.A.zero_rows_columns(self.bc, 1.0)
        if kkt:
            self.Bt, self.B, self.C, self.Q = da.assemble_kkt()
            self.Bt.zero_rows(self.bc, 0.0)
            self.B.zero_columns(self.bc)
            self.K = Mat.nest(self.A, self.Bt, self.B, self.C)
        elif constraints:
            self.B, self.Bt = da.assemble_constraints()
            self.Bt.zero_rows(self.bc, 0.0)
            self.B.zero_columns(self.bc)
            self.C = self.Q = None
            self.K = Mat.nest(self.A, self.Bt, self.B, None)
            gi = np.arange(self.nu, self.nu + 4, dtype=np.int32)       # AssembleRHS_Constraints: g
            self.rhs.set_values(gi, np.asarray(g, dtype=np.float64))
            self.kkt = True
        else:
            self.K = self.A
        self.n = n

    def make_ksp(self, options):
        ksp = KSP(self.ctx)
        ksp.set_operators(self.K, self.K)
        if self.kkt and self.Q is not None:
            ksp.set_schur_user_mat(self.Q)
        ksp.set_dmda(self.da)
        ksp.set_options(options)
        return ksp


class SaddlePointProblem3D:
    """The 3-D Stokes-type KKT problem of BASELINE config 4: [A Bt; B C] on a Q1-hexahedron grid, velocity Dirichlet on the
    whole boundary, rotational body force; Q (= -pressure mass) is the `user` Schur preconditioning matrix."""

    def __init__(self, ctx, nx, ny, nz, rhs_kind=1):
        self.ctx = ctx
        self.da = DMDA3D(ctx, nx, ny, nz)
        da = self.da
        self.nu, self.np_ = 3 * da.n_nodes_local, da.n_nodes_local
        self.n = self.nu + self.np_
        self.A = da.assemble_stress()
        self.rhs = Vec(ctx, self.n)
        da.assemble_rhs(self.rhs, rhs_kind)
        self.bc = da.bc_ids(3)
        self.rhs.set_values(self.bc, np.zeros(len(self.bc)))
        self.A.zero_rows_columns(self.bc, 1.0)
        self.Bt, self.B, self.C, self.Q = da.assemble_kkt()
        self.Bt.zero_rows(self.bc, 0.0)
        self.B.zero_columns(self.bc)
        self.K = Mat.nest(self.A, self.Bt, self.B, self.C)
        self.kkt = True

    def make_ksp(self, options):
        ksp = KSP(self.ctx)
        ksp.set_operators(self.K, self.K)
        ksp.set_schur_user_mat(self.Q)
        ksp.set_options(options)
        return ksp
