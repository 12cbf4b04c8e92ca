"""ctypes front-end of the CPU oracle (oracle/clipper_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(clipper_b200) never imports this module.

Array conventions follow the reference (Eigen, column-major):
  D1, D2 : numpy float64, shape (d, n)  -> passed in Fortran order (each datum contiguous)
  A      : numpy int32,   shape (m, 2)  -> passed in Fortran order (A[:,0] then A[:,1])
  M, C   : numpy float64, shape (m, m), symmetric, unit diagonal (getters), any order.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

NONZERO, DSD, DSD_HEU = 0, 1, 2


def build(force=False):
    """Compile liboracle.so with the committed Makefile (gcc, no other deps)."""
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "clipper_oracle.c"))
    ):
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], env=env,
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


class Params(C.Structure):
    """Mirror of clipper::Params (reference include/clipper/clipper.h:27-60)."""
    _fields_ = [
        ("tol_u", C.c_double), ("tol_F", C.c_double), ("tol_Fop", C.c_double),
        ("maxiniters", C.c_int32), ("maxoliters", C.c_int32),
        ("beta", C.c_double), ("maxlsiters", C.c_int32),
        ("eps", C.c_double), ("affinityeps", C.c_double),
        ("rescale_u0", C.c_int32), ("rounding", C.c_int32),
    ]


class _Solution(C.Structure):
    _fields_ = [
        ("t", C.c_double), ("ifinal", C.c_int32), ("n_nodes", C.c_int32),
        ("score", C.c_double), ("d_final", C.c_double),
        ("n_evals", C.c_int64), ("n_spmv", C.c_int64), ("n_inner", C.c_int64),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, dp, ip, lp = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
        L.orc_create.restype = vp
        L.orc_destroy.argtypes = [vp]
        L.orc_m.restype = C.c_int64; L.orc_m.argtypes = [vp]
        L.orc_nnz.restype = C.c_int64; L.orc_nnz.argtypes = [vp, C.c_int]
        L.orc_default_params.argtypes = [C.POINTER(Params)]
        L.orc_k2ij.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.orc_create_all_to_all.argtypes = [C.c_int64, C.c_int64, ip]
        L.orc_find_above.restype = C.c_int32
        L.orc_find_above.argtypes = [dp, C.c_int64, C.c_double, ip]
        L.orc_find_k_largest.restype = C.c_int32
        L.orc_find_k_largest.argtypes = [dp, C.c_int64, C.c_int32, ip]
        L.orc_euclidean.restype = C.c_double
        L.orc_euclidean.argtypes = [dp, dp, dp, dp, C.c_int, C.c_double, C.c_double, C.c_double]
        L.orc_pointnormal.restype = C.c_double
        L.orc_pointnormal.argtypes = [dp, dp, dp, dp, C.c_double, C.c_double, C.c_double, C.c_double]
        L.orc_score_euclidean.argtypes = [vp, dp, C.c_int32, C.c_int64, dp, C.c_int64, ip, C.c_int64,
                                          C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]
        L.orc_score_pointnormal.argtypes = [vp, dp, C.c_int64, dp, C.c_int64, ip, C.c_int64,
                                            C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]
        L.orc_set_dense.argtypes = [vp, dp, dp, C.c_int64]
        L.orc_set_sparse_upper.argtypes = [vp, C.c_int64, lp, ip, dp, lp, ip, dp]
        L.orc_get_csc.argtypes = [vp, C.c_int, lp, ip, dp]
        L.orc_get_dense.argtypes = [vp, C.c_int, dp]
        L.orc_get_associations.argtypes = [vp, ip]
        L.orc_matvec.argtypes = [vp, C.c_int, dp, dp]
        L.orc_gradf.argtypes = [vp, dp, C.c_double, dp, dp]
        L.orc_dsd_dense.restype = C.c_int32
        L.orc_dsd_dense.argtypes = [dp, C.c_int64, ip, C.c_int32, ip]
        L.orc_solve.argtypes = [vp, dp, C.POINTER(Params), C.POINTER(_Solution), dp, ip, dp, C.c_int64]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _l(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def default_params(**kw):
    p = Params()
    lib().orc_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def k2ij(k, n):
    i, j = C.c_uint64(), C.c_uint64()
    lib().orc_k2ij(k, n, C.byref(i), C.byref(j))
    return int(i.value), int(j.value)


def create_all_to_all(n1, n2):
    A = np.zeros((n1 * n2, 2), dtype=np.int32, order="F")
    lib().orc_create_all_to_all(n1, n2, _i(A))
    return A


def find_k_largest(x, k):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.zeros(max(int(k), 1) if k <= x.size else x.size, dtype=np.int32)
    n = lib().orc_find_k_largest(_d(x), x.size, int(k), _i(out))
    return out[:n].copy()


def find_above(x, thr):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.zeros(max(x.size, 1), dtype=np.int32)
    n = lib().orc_find_above(_d(x), x.size, float(thr), _i(out))
    return out[:n].copy()


def euclidean(ai, aj, bi, bj, sigma=0.01, epsilon=0.06, mindist=0.0):
    v = [np.ascontiguousarray(x, dtype=np.float64) for x in (ai, aj, bi, bj)]
    return lib().orc_euclidean(_d(v[0]), _d(v[1]), _d(v[2]), _d(v[3]), v[0].size, sigma, epsilon, mindist)


def pointnormal(ai, aj, bi, bj, sigp=0.5, epsp=0.5, sign=0.10, epsn=0.35):
    v = [np.ascontiguousarray(x, dtype=np.float64) for x in (ai, aj, bi, bj)]
    return lib().orc_pointnormal(_d(v[0]), _d(v[1]), _d(v[2]), _d(v[3]), sigp, epsp, sign, epsn)


def dsd_solve(A, S=()):
    """reference dsd::solve(const Eigen::MatrixXd&, S) (src/dsd.cpp:322-325): A dense symmetric (the strict upper
    triangle is what is read), S the node subset (empty: all nodes).  Returns the selected nodes, ascending."""
    A = np.asfortranarray(A, dtype=np.float64)
    n = A.shape[0]
    S = np.ascontiguousarray(np.asarray(S, dtype=np.int32))
    out = np.zeros(max(n, 1), np.int32)
    k = lib().orc_dsd_dense(_d(A), n, _i(S) if S.size else None, int(S.size), _i(out))
    return out[:k].tolist()


class Solution:
    def __init__(self):
        self.t = 0.0; self.ifinal = 0; self.nodes = np.zeros(0, np.int32)
        self.u0 = None; self.u = None; self.score = 0.0
        self.d_final = 0.0; self.n_evals = 0; self.n_spmv = 0; self.n_inner = 0
        self.trace = None; self.dsd_support = None


class Oracle:
    """CPU mirror of clipper::CLIPPER for the hot path (reference clipper.h:78-148)."""

    def __init__(self, params=None):
        self.params = params if params is not None else default_params()
        self._h = C.c_void_p(lib().orc_create())
        self.soln = Solution()

    def __del__(self):
        try:
            if self._h:
                lib().orc_destroy(self._h); self._h = None
        except Exception:
            pass

    @property
    def m(self):
        return int(lib().orc_m(self._h))

    def nnz(self, which=0):
        return int(lib().orc_nnz(self._h, which))

    def score_euclidean(self, D1, D2, A=None, sigma=0.01, epsilon=0.06, mindist=0.0, nthreads=0):
        D1 = np.asfortranarray(D1, dtype=np.float64); D2 = np.asfortranarray(D2, dtype=np.float64)
        d, n1 = D1.shape; n2 = D2.shape[1]
        if A is None or np.size(A) == 0:
            Ap, m = None, 0
        else:
            A = np.asfortranarray(A, dtype=np.int32); Ap, m = _i(A), A.shape[0]
        lib().orc_score_euclidean(self._h, _d(D1), d, n1, _d(D2), n2, Ap, m, sigma, epsilon, mindist,
                                  self.params.affinityeps, nthreads)

    def score_pointnormal(self, D1, D2, A=None, sigp=0.5, epsp=0.5, sign=0.10, epsn=0.35, nthreads=0):
        D1 = np.asfortranarray(D1, dtype=np.float64); D2 = np.asfortranarray(D2, dtype=np.float64)
        assert D1.shape[0] == 6 and D2.shape[0] == 6
        n1, n2 = D1.shape[1], D2.shape[1]
        if A is None or np.size(A) == 0:
            Ap, m = None, 0
        else:
            A = np.asfortranarray(A, dtype=np.int32); Ap, m = _i(A), A.shape[0]
        lib().orc_score_pointnormal(self._h, _d(D1), n1, _d(D2), n2, Ap, m, sigp, epsp, sign, epsn,
                                    self.params.affinityeps, nthreads)

    def set_matrix_data(self, M, Cm):
        M = np.asfortranarray(M, dtype=np.float64); Cm = np.asfortranarray(Cm, dtype=np.float64)
        lib().orc_set_dense(self._h, _d(M), _d(Cm), M.shape[0])

    def set_sparse_upper(self, m, cpM, riM, vM, cpC, riC, vC):
        a = [np.ascontiguousarray(cpM, np.int64), np.ascontiguousarray(riM, np.int32),
             np.ascontiguousarray(vM, np.float64), np.ascontiguousarray(cpC, np.int64),
             np.ascontiguousarray(riC, np.int32), np.ascontiguousarray(vC, np.float64)]
        lib().orc_set_sparse_upper(self._h, m, _l(a[0]), _i(a[1]), _d(a[2]), _l(a[3]), _i(a[4]), _d(a[5]))

    def get_csc(self, which=0):
        m, nnz = self.m, self.nnz(which)
        cp = np.zeros(m + 1, np.int64); ri = np.zeros(max(nnz, 1), np.int32); v = np.zeros(max(nnz, 1), np.float64)
        lib().orc_get_csc(self._h, which, _l(cp), _i(ri), _d(v))
        return cp, ri[:nnz], v[:nnz]

    def get_affinity_matrix(self):
        m = self.m; out = np.zeros((m, m), dtype=np.float64, order="F")
        lib().orc_get_dense(self._h, 0, _d(out)); return out

    def get_constraint_matrix(self):
        m = self.m; out = np.zeros((m, m), dtype=np.float64, order="F")
        lib().orc_get_dense(self._h, 1, _d(out)); return out

    def get_initial_associations(self):
        A = np.zeros((self.m, 2), dtype=np.int32, order="F")
        lib().orc_get_associations(self._h, _i(A)); return A

    def matvec(self, x, which=0):
        x = np.ascontiguousarray(x, dtype=np.float64); y = np.zeros_like(x)
        lib().orc_matvec(self._h, which, _d(x), _d(y)); return y

    def gradf(self, v, d):
        v = np.ascontiguousarray(v, dtype=np.float64); y = np.zeros_like(v); F = C.c_double()
        lib().orc_gradf(self._h, _d(v), float(d), _d(y), C.byref(F)); return y, F.value

    def solve(self, u0, trace_cap=0):
        """u0 must be explicit: the reference default is seeded from std::random_device (utils.cpp:24)."""
        m = self.m
        u0 = np.ascontiguousarray(u0, dtype=np.float64); assert u0.size == m
        u = np.zeros(m, np.float64); nodes = np.zeros(max(m, 1), np.int32)
        tr = np.zeros((max(trace_cap, 1), 3), np.float64)
        s = _Solution()
        lib().orc_solve(self._h, _d(u0), C.byref(self.params), C.byref(s), _d(u), _i(nodes),
                        _d(tr) if trace_cap else None, trace_cap)
        out = Solution()
        out.t, out.ifinal, out.score, out.d_final = s.t, s.ifinal, s.score, s.d_final
        out.n_evals, out.n_spmv, out.n_inner = s.n_evals, s.n_spmv, s.n_inner
        out.u0, out.u = u0.copy(), u
        out.nodes = nodes[: s.n_nodes].copy()
        out.trace = tr[: min(trace_cap, s.ifinal + 1)] if trace_cap else None
        self.soln = out
        return out

    def get_selected_associations(self):
        A = self.get_initial_associations()
        return A[self.soln.nodes, :]
