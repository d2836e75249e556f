"""ctypes binding + solver composition for the CPU ORACLE (test infrastructure, NOT the product).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
The C restatement lives in sp_oracle.c; this file wires PETSc-style option dictionaries
 to the oracle's operator / KSP objects.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


class CsrStruct(C.Structure):
    _fields_ = [("nrows", C.c_int), ("ncols", C.c_int), ("rowptr", c_ip), ("col", c_ip), ("val", c_dp)]


class KspStruct(C.Structure):
    _fields_ = [("type", C.c_int), ("A", C.c_void_p), ("M", C.c_void_p), ("rtol", C.c_double), ("atol", C.c_double),
                ("dtol", C.c_double), ("max_it", C.c_int), ("restart", C.c_int), ("norm_none", C.c_int),
                ("emin", C.c_double), ("emax", C.c_double), ("richardson_scale", C.c_double), ("orthog", C.c_int), ("its", C.c_int),
                ("reason", C.c_int), ("rnorm", C.c_double), ("rnorm0", C.c_double), ("hist", c_dp),
                ("hist_cap", C.c_int), ("hist_len", C.c_int)]


CsrP = C.POINTER(CsrStruct)
KspP = C.POINTER(KspStruct)


def build(force=False):
    so = os.path.join(_HERE, "libsp_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("sp_oracle.c", "sp_oracle3d.c", "sp_oracle_amg.c", "sp_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.exists(f) and os.path.getmtime(f) > os.path.getmtime(so) for f in srcs):
        subprocess.check_call(["make", "-C", _HERE, "libsp_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    L = C.CDLL(build())
    vp = C.c_void_p

    def sig(name, res, *args):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = list(args)

    sig("or_csr_alloc", CsrP, C.c_int, C.c_int, C.c_long)
    sig("or_csr_free", None, CsrP)
    sig("or_csr_nnz", C.c_long, CsrP)
    sig("or_csr_transpose", CsrP, CsrP)
    sig("or_csr_matmat", CsrP, CsrP, CsrP)
    sig("or_csr_add_scaled", CsrP, CsrP, C.c_double, CsrP)
    sig("or_csr_scale_cols", CsrP, CsrP, c_dp)
    sig("or_csr_get_diagonal", None, CsrP, c_dp)
    sig("or_csr_mult", None, CsrP, c_dp, c_dp)
    sig("or_csr_mult_add", None, CsrP, c_dp, c_dp)
    sig("or_dmda_proc_grid", None, C.c_int, C.c_int, C.c_int, c_ip, c_ip)
    sig("or_dmda_ownership", None, C.c_int, C.c_int, c_ip)
    sig("or_dmda_natural_to_petsc", None, C.c_int, C.c_int, C.c_int, c_ip, c_ip)
    sig("or_dmda_element_range", None, C.c_int, C.c_int, C.c_int, C.c_int, c_ip, c_ip, c_ip, c_ip)
    sig("or_element_coords", None, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_dp)
    sig("or_element_stress", None, c_dp, c_dp, c_dp)
    sig("or_element_rhs", None, c_dp, C.c_int, c_dp)
    sig("or_element_kkt", None, c_dp, c_dp, c_dp, c_dp)
    sig("or_assemble_A", CsrP, C.c_int, C.c_int, C.c_int)
    sig("or_assemble_A_coeff", CsrP, C.c_int, C.c_int, C.c_int)
    sig("or_assemble_rhs", None, C.c_int, C.c_int, C.c_int, C.c_int, c_dp)
    sig("or_bc_ids", C.c_int, C.c_int, C.c_int, C.c_int, c_ip)
    sig("or_apply_bc", None, CsrP, c_dp, C.c_int, c_ip)
    sig("or_assemble_kkt", None, C.c_int, C.c_int, C.POINTER(CsrP), C.POINTER(CsrP), C.POINTER(CsrP), C.POINTER(CsrP))
    sig("or3_assemble_A", CsrP, C.c_int, C.c_int, C.c_int)
    sig("or3_assemble_rhs", None, C.c_int, C.c_int, C.c_int, C.c_int, c_dp)
    sig("or3_assemble_kkt", None, C.c_int, C.c_int, C.c_int, C.POINTER(CsrP), C.POINTER(CsrP), C.POINTER(CsrP), C.POINTER(CsrP))
    sig("or3_bc_ids", C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_ip)
    sig("or3_dmda_proc_grid", None, C.c_int, C.c_int, C.c_int, C.c_int, c_ip, c_ip, c_ip)
    sig("or3_dmda_natural_to_petsc", None, C.c_int, C.c_int, C.c_int, C.c_int, c_ip, c_ip)
    sig("or3_element_coords", None, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_dp)
    sig("or3_element_stress", None, c_dp, c_dp)
    sig("or3_element_rhs", None, c_dp, C.c_int, c_dp)
    sig("or3_element_kkt", None, c_dp, c_dp, c_dp, c_dp)
    sig("or_assemble_constraints", None, C.c_int, C.c_int, C.POINTER(CsrP), C.POINTER(CsrP))
    sig("or_element_constraints", None, c_dp, c_dp)
    sig("or_zero_rows", None, CsrP, C.c_int, c_ip)
    sig("or_zero_cols", None, CsrP, C.c_int, c_ip)
    sig("or_amg_aggregate", C.c_int, CsrP, C.c_int, C.c_double, C.c_int, c_ip)
    sig("or_amg_tentative", CsrP, C.c_int, C.c_int, c_ip, C.c_int, c_ip, c_ip)
    sig("or_csr_scale_rows", CsrP, CsrP, c_dp)
    sig("or_amg_smooth_prolongator", CsrP, CsrP, CsrP, C.c_double)
    sig("or_amg_galerkin", CsrP, CsrP, CsrP)
    sig("or_interp_q1", CsrP, C.c_int, C.c_int, C.c_int, C.c_int)
    sig("or_op_apply", None, vp, c_dp, c_dp)
    sig("or_op_free", None, vp)
    sig("or_op_csr", vp, CsrP)
    sig("or_op_jacobi", vp, CsrP)
    sig("or_op_diag_inverse", vp, C.c_int, c_dp)
    sig("or_op_nest", vp, CsrP, CsrP, CsrP, CsrP)
    sig("or_op_schur", vp, CsrP, CsrP, vp, CsrP)
    sig("or_op_fieldsplit", vp, C.c_int, CsrP, CsrP, vp, vp, C.c_double)
    sig("or_op_lsc", vp, CsrP, CsrP, CsrP, vp, C.c_int)
    sig("or_op_permuted", vp, vp, C.c_int, c_ip)
    sig("or_op_dense_lu", vp, CsrP)
    sig("or_op_mg", vp, C.c_int, C.POINTER(CsrP), C.POINTER(CsrP), C.POINTER(KspP), vp)
    sig("or_ksp_create", KspP, C.c_int, vp, vp)
    sig("or_ksp_free", None, KspP)
    sig("or_ksp_set_history", None, KspP, C.c_int)
    sig("or_ksp_solve", C.c_int, KspP, c_dp, c_dp, C.c_int)
    sig("or_op_from_ksp", vp, KspP)
    sig("or_estimate_lambda_max", C.c_double, vp, vp, C.c_int)
    sig("or_hash_vector", None, C.c_int, c_dp)
    sig("or_dot", C.c_double, C.c_int, c_dp, c_dp)
    sig("or_norm2", C.c_double, C.c_int, c_dp)
    sig("or_set_threads", None, C.c_int)
    sig("or_get_threads", C.c_int)
    _LIB = L
    return L


def dptr(a):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(c_dp)


def iptr(a):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(c_ip)


class Csr:
    """Owning wrapper of an OrCsr* with numpy views."""

    def __init__(self, ptr, own=True):
        self.ptr = ptr
        self.own = own
        s = ptr.contents
        self.nrows, self.ncols = s.nrows, s.ncols
        self.rowptr = np.ctypeslib.as_array(s.rowptr, shape=(self.nrows + 1,))
        nnz = int(self.rowptr[-1])
        self.nnz = nnz
        self.col = np.ctypeslib.as_array(s.col, shape=(max(nnz, 1),))[:nnz]
        self.val = np.ctypeslib.as_array(s.val, shape=(max(nnz, 1),))[:nnz]

    @classmethod
    def from_arrays(cls, nrows, ncols, rowptr, col, val):
        p = lib().or_csr_alloc(nrows, ncols, len(col))
        cls(p, own=False).rowptr[:] = rowptr   # the views of an OrCsr are sized from rowptr: fill it first
        c = cls(p)
        c.col[:] = col
        c.val[:] = val
        return c

    def scipy(self):
        import scipy.sparse as sp
        return sp.csr_matrix((self.val.copy(), self.col.copy(), self.rowptr.copy()), shape=(self.nrows, self.ncols))

    def mult(self, x):
        y = np.empty(self.nrows)
        lib().or_csr_mult(self.ptr, dptr(np.ascontiguousarray(x, dtype=np.float64)), dptr(y))
        return y

    def diagonal(self):
        d = np.empty(self.nrows)
        lib().or_csr_get_diagonal(self.ptr, dptr(d))
        return d

    def transpose(self):
        return Csr(lib().or_csr_transpose(self.ptr))

    def matmat(self, other):
        return Csr(lib().or_csr_matmat(self.ptr, other.ptr))

    def __del__(self):
        if getattr(self, "own", False) and self.ptr and _LIB is not None:
            _LIB.or_csr_free(self.ptr)
            self.ptr = None


# ------------------------------------------------------------------ problems
def dmda_proc_grid(M, N, size):
    m, n = C.c_int(), C.c_int()
    lib().or_dmda_proc_grid(M, N, size, C.byref(m), C.byref(n))
    return m.value, n.value


def dmda_ownership(M, m):
    lx = np.zeros(m, dtype=np.int32)
    lib().or_dmda_ownership(M, m, iptr(lx))
    return lx


def dmda_natural_to_petsc(M, N, size):
    nm = np.zeros(M * N, dtype=np.int32)
    ow = np.zeros(M * N, dtype=np.int32)
    lib().or_dmda_natural_to_petsc(M, N, size, iptr(nm), iptr(ow))
    return nm, ow


def dmda3d_natural_to_petsc(M, N, P, size):
    nm = np.zeros(M * N * P, dtype=np.int32)
    ow = np.zeros(M * N * P, dtype=np.int32)
    lib().or3_dmda_natural_to_petsc(M, N, P, size, iptr(nm), iptr(ow))
    return nm, ow


def dmda_element_range(M, N, size, rank):
    v = [C.c_int() for _ in range(4)]
    lib().or_dmda_element_range(M, N, size, rank, *[C.byref(x) for x in v])
    return tuple(x.value for x in v)


def element_stress(ec, coeff=None):
    ke = np.zeros(64)
    co = np.ones(4) if coeff is None else np.asarray(coeff, dtype=np.float64)
    lib().or_element_stress(dptr(np.ascontiguousarray(ec, dtype=np.float64)), dptr(co), dptr(ke))
    return ke


def element_rhs(ec, kind=0):
    fe = np.zeros(8)
    lib().or_element_rhs(dptr(np.ascontiguousarray(ec, dtype=np.float64)), kind, dptr(fe))
    return fe


def element_coords(M, N, ei, ej, as_written=False):
    ec = np.zeros(8)
    lib().or_element_coords(M, N, ei, ej, int(as_written), dptr(ec))
    return ec


def bc_ids(M, N, dof=2):
    ids = np.zeros(dof * (2 * M + 2 * N - 4), dtype=np.int32)
    n = lib().or_bc_ids(M, N, dof, iptr(ids))
    assert n == len(ids)
    return ids


class Problem:
    """The reference's A u = f (velocity block) and, with kkt=True, the [A Bt; B C] extension."""

    def __init__(self, nx, ny, kkt=False, rhs_kind=0, as_written=False, bc=True, constraints=False, g=(0.0, 0.0, 0.0, 0.0)):
        L = lib()
        self.M, self.N = nx + 1, ny + 1
        M, N = self.M, self.N
        self.nu, self.np_ = 2 * M * N, M * N
        self.A = Csr(L.or_assemble_A(M, N, int(as_written)))
        self.f = np.zeros(self.nu)
        L.or_assemble_rhs(M, N, int(as_written), rhs_kind, dptr(self.f))
        self.bc = bc_ids(M, N, 2)
        if bc:
            L.or_apply_bc(self.A.ptr, dptr(self.f), len(self.bc), iptr(self.bc))
        self.kkt = kkt
        if kkt:
            ps = [CsrP() for _ in range(4)]
            L.or_assemble_kkt(M, N, *[C.byref(p) for p in ps])
            self.Bt, self.B, self.C, self.Q = [Csr(p) for p in ps]
            if bc:
                L.or_zero_rows(self.Bt.ptr, len(self.bc), iptr(self.bc))
                L.or_zero_cols(self.B.ptr, len(self.bc), iptr(self.bc))
            self.rhs = np.concatenate([self.f, np.zeros(self.np_)])
        elif constraints:
            # the reference's own saddle-point problem (src/SaddlePointProblem.c:45-60, commented out there):
            # [A Bt; B 0] [u; lambda] = [f; g] with the 4 dense barycentre / moment rows of or_assemble_constraints
            ps = [CsrP(), CsrP()]
            L.or_assemble_constraints(M, N, C.byref(ps[0]), C.byref(ps[1]))
            self.B, self.Bt = Csr(ps[0]), Csr(ps[1])
            if bc:
                L.or_zero_rows(self.Bt.ptr, len(self.bc), iptr(self.bc))
                L.or_zero_cols(self.B.ptr, len(self.bc), iptr(self.bc))
            self.C = self.Q = None
            self.np_ = 4
            self.kkt = True
            self.rhs = np.concatenate([self.f, np.asarray(g, dtype=np.float64)])
        else:
            self.rhs = self.f

    def operator(self):
        if self.kkt:
            return lib().or_op_nest(self.A.ptr, self.Bt.ptr, self.B.ptr, self.C.ptr if self.C is not None else None)
        return lib().or_op_csr(self.A.ptr)

    def scipy_K(self):
        import scipy.sparse as sp
        if not self.kkt:
            return self.A.scipy()
        return sp.bmat([[self.A.scipy(), self.Bt.scipy()], [self.B.scipy(), self.C.scipy() if self.C is not None else None]], format="csr")


class Problem3D:
    """3-D Q1-hexahedron KKT problem [A Bt; B C] on (nx+1)(ny+1)(nz+1) nodes (BASELINE config 4; sp_oracle3d.c)."""

    def __init__(self, nx, ny, nz, rhs_kind=1, bc=True):
        L = lib()
        self.M, self.N, self.P = nx + 1, ny + 1, nz + 1
        M, N, P = self.M, self.N, self.P
        nn = M * N * P
        self.nu, self.np_ = 3 * nn, nn
        self.A = Csr(L.or3_assemble_A(M, N, P))
        self.f = np.zeros(self.nu)
        L.or3_assemble_rhs(M, N, P, rhs_kind, dptr(self.f))
        nbc = L.or3_bc_ids(M, N, P, 3, None)
        self.bc = np.zeros(nbc, dtype=np.int32)
        L.or3_bc_ids(M, N, P, 3, iptr(self.bc))
        if bc:
            L.or_apply_bc(self.A.ptr, dptr(self.f), len(self.bc), iptr(self.bc))
        ps = [CsrP() for _ in range(4)]
        L.or3_assemble_kkt(M, N, P, *[C.byref(p) for p in ps])
        self.Bt, self.B, self.C, self.Q = [Csr(p) for p in ps]
        if bc:
            L.or_zero_rows(self.Bt.ptr, len(self.bc), iptr(self.bc))
            L.or_zero_cols(self.B.ptr, len(self.bc), iptr(self.bc))
        self.kkt = True
        self.rhs = np.concatenate([self.f, np.zeros(self.np_)])

    def operator(self):
        return lib().or_op_nest(self.A.ptr, self.Bt.ptr, self.B.ptr, self.C.ptr)

    def scipy_K(self):
        import scipy.sparse as sp
        return sp.bmat([[self.A.scipy(), self.Bt.scipy()], [self.B.scipy(), self.C.scipy()]], format="csr")


MIS_ORDER = {"hash": 0, "natural": 1}


def amg_aggregate(mat, bs, theta=0.0, order="hash"):
    """(aggregate id per node, -1 = left out; number of aggregates) -- or_amg_aggregate."""
    nn = mat.nrows // bs
    agg = np.zeros(max(nn, 1), dtype=np.int32)
    nagg = lib().or_amg_aggregate(mat.ptr, bs, float(theta), MIS_ORDER[order], iptr(agg))
    return agg[:nn], nagg


def amg_hierarchy(mat, bs, theta=0.0, nsmooths=1, coarse_limit=50, max_levels=30, order="hash"):
    """Level matrices [A_0 .. A_L], prolongators [P_0 .. P_{L-1}] (P_l: level l+1 -> l) and the aggregates per level."""
    L = lib()
    mats, interps, aggs = [mat], [], []
    w = None   # finest-level nodes behind every node of the current level (None: ones)
    while len(mats) < max_levels and mats[-1].nrows > coarse_limit:
        A = mats[-1]
        n = A.nrows
        agg, nagg = amg_aggregate(A, bs, theta, order)
        if nagg == 0 or nagg * bs >= n:
            break
        wc = np.zeros(nagg, dtype=np.int32)
        Pt = Csr(L.or_amg_tentative(n // bs, bs, iptr(agg), nagg, iptr(w) if w is not None else None, iptr(wc)))
        if nsmooths:
            Aop = L.or_op_csr(A.ptr)
            Jop = L.or_op_jacobi(A.ptr)
            lam = L.or_estimate_lambda_max(Aop, Jop, 10)
            L.or_op_free(Aop); L.or_op_free(Jop)
            P = Csr(L.or_amg_smooth_prolongator(A.ptr, Pt.ptr, 4.0 / (3.0 * lam)))
        else:
            P = Pt
        interps.append(P)
        aggs.append((agg, nagg, w))
        w = wc
        mats.append(Csr(L.or_amg_galerkin(A.ptr, P.ptr)))
    return mats, interps, aggs


# ------------------------------------------------------------- option wiring
KSP_TYPES = {"preonly": 0, "richardson": 1, "chebyshev": 2, "gmres": 3, "fgmres": 4, "minres": 5}
FACT = {"diag": 0, "lower": 1, "upper": 2, "full": 3}


def parse_options(s):
    """'-ksp_type gmres -ksp_rtol 1e-8 -flag' -> {'ksp_type': 'gmres', 'ksp_rtol': '1e-8', 'flag': ''}"""
    if isinstance(s, dict):
        return dict(s)
    toks = s.split()
    out = {}
    i = 0
    while i < len(toks):
        t = toks[i]
        assert t.startswith("-"), t
        key = t[1:]
        if i + 1 < len(toks) and not (toks[i + 1].startswith("-") and not _is_number(toks[i + 1])):
            out[key] = toks[i + 1]
            i += 2
        else:
            out[key] = ""
            i += 1
    return out


def _is_number(t):
    try:
        float(t)
        return True
    except ValueError:
        return False


class Solver:
    """Composes oracle objects from PETSc-style options and keeps them alive."""

    def __init__(self, prob, options):
        self.L = lib()
        self.prob = prob
        self.o = parse_options(options)
        self.keep = []
        self.ksps = {}
        self.Aop = prob.operator()
        self.keep.append(self.Aop)
        pc = self._make_pc("", prob)
        self.ksp = self._make_ksp("", self.Aop, pc, default_type="gmres")

    # -- helpers
    def _get(self, key, default=None):
        return self.o.get(key, default)

    def _make_ksp(self, prefix, Aop, Mop, default_type="preonly", default_max_it=10000):
        L = self.L
        t = self._get(prefix + "ksp_type", default_type)
        k = L.or_ksp_create(KSP_TYPES[t], Aop, Mop)
        s = k.contents
        s.rtol = float(self._get(prefix + "ksp_rtol", 1e-5))
        s.atol = float(self._get(prefix + "ksp_atol", 1e-50))
        s.max_it = int(self._get(prefix + "ksp_max_it", default_max_it))
        s.restart = int(self._get(prefix + "ksp_gmres_restart", 30))
        s.richardson_scale = float(self._get(prefix + "ksp_richardson_scale", 1.0))
        if self._get(prefix + "ksp_gmres_modifiedgramschmidt") is not None:
            s.orthog = 3
        else:
            s.orthog = {"refine_never": 0, "refine_ifneeded": 1, "refine_always": 2}[self._get(prefix + "ksp_gmres_cgs_refinement_type", "refine_never")]
        if self._get(prefix + "ksp_norm_type", "default") == "none" or (t in ("chebyshev", "richardson") and prefix and
                                                                      self._get(prefix + "ksp_norm_type") is None):
            s.norm_none = 1  # inner chebyshev/richardson are fixed-sweep smoothers unless a norm is requested
        if t == "chebyshev":
            ev = self._get(prefix + "ksp_chebyshev_eigenvalues")
            if ev:
                s.emin, s.emax = [float(x) for x in ev.split(",")]
            else:
                lam = L.or_estimate_lambda_max(Aop, Mop, 10)
                s.emin, s.emax = 0.1 * lam, 1.1 * lam
        self.ksps[prefix] = k
        return k

    def _ksp_op(self, k):
        op = self.L.or_op_from_ksp(k)
        self.keep.append(op)
        return op

    def _make_simple_pc(self, prefix, mat, grid=None, dof=1, default="none"):
        """pc on an assembled matrix: none | jacobi | mg | lu"""
        L = self.L
        t = self._get(prefix + "pc_type", default)
        if t == "petsc-default":   # PETSc would pick ILU(0) here; neither the oracle nor the CUDA library has it
            raise ValueError("oracle: -%spc_type must be given explicitly (PETSc's default is ILU(0))" % prefix)
        if t == "none":
            return None
        if t == "jacobi":
            op = L.or_op_jacobi(mat.ptr)
        elif t == "lu":
            op = L.or_op_dense_lu(mat.ptr)
        elif t == "mg":
            op = self._make_mg(prefix, mat, grid, dof)
        elif t == "gamg":
            op = self._make_gamg(prefix, mat, dof)
        else:
            raise ValueError("oracle: unsupported pc_type %r for %r" % (t, prefix))
        self.keep.append(op)
        return op

    def _make_mg(self, prefix, mat, grid, dof):
        """PCMG with rediscretised coarse operators (velocity block) and Q1 interpolation."""
        L = self.L
        M, N = grid
        nlev = int(self._get(prefix + "pc_mg_levels", 2))
        mats, interps, smooth = [mat], [], []
        Ml, Nl = M, N
        for l in range(1, nlev):
            assert (Ml - 1) % 2 == 0 and (Nl - 1) % 2 == 0, "grid not coarsenable"
            Mc, Nc = (Ml - 1) // 2 + 1, (Nl - 1) // 2 + 1
            if dof == 2:
                Ac = Csr(L.or_assemble_A(Mc, Nc, 0))
                ids = bc_ids(Mc, Nc, 2)
                L.or_apply_bc(Ac.ptr, None, len(ids), iptr(ids))
                interps.append(Csr(L.or_interp_q1(Mc, Nc, 2, 1)))
            else:
                raise ValueError("oracle mg: only the velocity block (dof 2) is rediscretised")
            mats.append(Ac)
            Ml, Nl = Mc, Nc
        return self._mg_from_levels(prefix, mats, interps)

    def _mg_from_levels(self, prefix, mats, interps):
        """or_op_mg on given level matrices / prolongators with the -<prefix>mg_levels_ smoothers and a dense coarse solve."""
        L = self.L
        nlev = len(mats)
        smooth = []
        self.keep += mats + interps
        for l in range(nlev - 1):
            Aop = L.or_op_csr(mats[l].ptr)
            Jop = L.or_op_jacobi(mats[l].ptr)
            self.keep += [Aop, Jop]
            sp_ = prefix + "mg_levels_"
            k = L.or_ksp_create(KSP_TYPES[self._get(sp_ + "ksp_type", "chebyshev")], Aop, Jop)
            k.contents.max_it = int(self._get(sp_ + "ksp_max_it", 2))
            k.contents.norm_none = 1
            lam = L.or_estimate_lambda_max(Aop, Jop, 10)
            k.contents.emin, k.contents.emax = 0.1 * lam, 1.1 * lam
            k.contents.richardson_scale = float(self._get(sp_ + "ksp_richardson_scale", 1.0))
            smooth.append(k)
        smooth.append(KspP())
        coarse = L.or_op_dense_lu(mats[-1].ptr)
        self.keep.append(coarse)
        self.mg_smooth = getattr(self, "mg_smooth", []) + smooth
        Aarr = (CsrP * nlev)(*[m.ptr for m in mats])
        Parr = (CsrP * nlev)(*([p.ptr for p in interps] + [CsrP()]))
        Sarr = (KspP * nlev)(*smooth)
        return L.or_op_mg(nlev, Aarr, Parr, Sarr, coarse)

    def _make_gamg(self, prefix, mat, bs):
        """Aggregation multigrid (sp_oracle_amg.c): levels by MIS-2 aggregation until n <= -pc_gamg_coarse_eq_limit."""
        mats, interps, self.gamg_aggregates = amg_hierarchy(
            mat, bs, theta=float(self._get(prefix + "pc_gamg_threshold", 0.0)),
            nsmooths=int(self._get(prefix + "pc_gamg_agg_nsmooths", 1)),
            coarse_limit=int(self._get(prefix + "pc_gamg_coarse_eq_limit", 50)),
            max_levels=int(self._get(prefix + "pc_mg_levels", 30)),
            order=self._get(prefix + "pc_gamg_mis_ordering", "hash"))
        return self._mg_from_levels(prefix, mats, interps)

    def _make_pc(self, prefix, prob):
        L = self.L
        t = self._get(prefix + "pc_type", "none")
        if t != "fieldsplit":
            return self._make_simple_pc(prefix, prob.A, grid=(prob.M, prob.N), dof=2)
        assert self._get("pc_fieldsplit_type", "schur") == "schur"
        if not prob.kkt:
            return self._make_strided_fieldsplit(prob)
        return self._make_pc_split(prob)

    def _make_pc_split(self, prob):
        L = self.L
        fact = FACT[self._get("pc_fieldsplit_schur_fact_type", "full")]
        pre = self._get("pc_fieldsplit_schur_precondition", "a11")
        scale = float(self._get("pc_fieldsplit_schur_scale", -1.0))
        # K0: fieldsplit_0 KSP on A00
        A00op = L.or_op_csr(prob.A.ptr)
        self.keep.append(A00op)
        pc0 = self._make_simple_pc("fieldsplit_0_", prob.A, grid=(prob.M, prob.N), dof=2, default="petsc-default")
        k0 = self._make_ksp("fieldsplit_0_", A00op, pc0, default_type="preonly")
        K0 = self._ksp_op(k0)
        # S = A11 - A10 ksp(A00) A01, with its own (identically configured) inner KSP
        k0s = self._make_ksp("fieldsplit_0_", A00op, pc0, default_type="preonly")
        self.ksps["schur_inner_"] = k0s
        K0s = self._ksp_op(k0s)
        Cptr = prob.C.ptr if prob.C is not None else None
        S = L.or_op_schur(Cptr, prob.B.ptr, K0s, prob.Bt.ptr)
        self.keep.append(S)
        # preconditioning matrix for the S solve
        if pre == "a11":
            Sp = prob.C
        elif pre == "user":
            Sp = prob.Q
        elif pre == "selfp":
            d = prob.A.diagonal()
            BtD = Csr(L.or_csr_scale_cols(prob.B.ptr, dptr(np.ascontiguousarray(1.0 / d))))  # A10 * D^-1
            prod = BtD.matmat(prob.Bt)
            if prob.C is not None:
                Sp = Csr(L.or_csr_add_scaled(prob.C.ptr, -1.0, prod.ptr))
            else:   # no (1,1) block: Sp = -A10 D^-1 A01 = prod + (-2) prod, exact in floating point
                Sp = Csr(L.or_csr_add_scaled(prod.ptr, -2.0, prod.ptr))
            self.keep += [BtD, prod]
        elif pre == "self":
            Sp = None
        else:
            raise ValueError(pre)
        self.Sp = Sp
        pt = self._get("fieldsplit_1_pc_type", "petsc-default" if Sp is not None else "none")
        if pt == "lsc":
            scale_diag = self._get("fieldsplit_1_pc_lsc_scale_diag") is not None
            if scale_diag:
                d = prob.A.diagonal()
                BD = Csr(L.or_csr_scale_cols(prob.B.ptr, dptr(np.ascontiguousarray(1.0 / d))))
                Lm = BD.matmat(prob.Bt)
                self.keep.append(BD)
            else:
                Lm = prob.B.matmat(prob.Bt)
            self.Lmat = Lm
            Lop = L.or_op_csr(Lm.ptr)
            self.keep += [Lm, Lop]
            pcl = self._make_simple_pc("fieldsplit_1_lsc_", Lm, default="petsc-default")
            kl = self._make_ksp("fieldsplit_1_lsc_", Lop, pcl, default_type="gmres")   # a fresh KSP in PETSc: GMRES
            Linv = self._ksp_op(kl)
            pcS = L.or_op_lsc(prob.A.ptr, prob.Bt.ptr, prob.B.ptr, Linv, int(scale_diag))
            self.keep.append(pcS)
        elif pt == "none" or Sp is None:
            pcS = None
        else:
            pcS = self._make_simple_pc("fieldsplit_1_", Sp, default="petsc-default")
        kS = self._make_ksp("fieldsplit_1_", S, pcS, default_type="gmres")   # the Schur KSP is a fresh KSP in PETSc: GMRES
        KS = self._ksp_op(kS)
        fs = L.or_op_fieldsplit(fact, prob.Bt.ptr, prob.B.ptr, K0, KS, scale)
        self.keep.append(fs)
        return fs

    def _make_strided_fieldsplit(self, prob):
        """PCFIELDSPLIT on the reference's own operator: KSPSetOperators(A,A) with block size 2 and no DM on the KSP
        -> split 0 = Ux dofs, split 1 = Uy dofs (strided ISs), sub-matrices by MatCreateSubMatrix."""
        import types
        bs = int(self._get("pc_fieldsplit_block_size", 2))
        assert bs == 2
        A = prob.A.scipy()
        n = A.shape[0]
        i0, i1 = np.arange(0, n, 2), np.arange(1, n, 2)

        def sub(r, c):
            S = A[r][:, c].tocsr()
            S.sort_indices()
            return Csr.from_arrays(S.shape[0], S.shape[1], S.indptr.astype(np.int32), S.indices.astype(np.int32), S.data)

        sp_ = types.SimpleNamespace(A=sub(i0, i0), Bt=sub(i0, i1), B=sub(i1, i0), C=sub(i1, i1), Q=None, M=prob.M, N=prob.N, kkt=True)
        self.keep.append(sp_)
        inner = self._make_pc_split(sp_)
        mp = np.concatenate([i0, i1]).astype(np.int32)
        self.keep.append(mp)
        op = self.L.or_op_permuted(inner, n, iptr(mp))
        self.keep.append(op)
        return op

    def solve(self, b=None, history=True):
        b = self.prob.rhs if b is None else b
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.zeros_like(b)
        if history:
            self.L.or_ksp_set_history(self.ksp, int(self.ksp.contents.max_it) + 2)
        reason = self.L.or_ksp_solve(self.ksp, dptr(b), dptr(x), 0)
        s = self.ksp.contents
        hist = np.array([s.hist[i] for i in range(s.hist_len)]) if history else None
        return {"x": x, "its": s.its, "reason": reason, "rnorm": s.rnorm, "rnorm0": s.rnorm0, "history": hist}

    def apply_operator(self, x):
        y = np.empty(len(x))
        self.L.or_op_apply(self.Aop, dptr(np.ascontiguousarray(x, dtype=np.float64)), dptr(y))
        return y
