"""Batches of small problems in one launch -- host mirror of clp_batch_* (include/clipper_b200.h).

The reference's own operating point is m <= 2048 associations per registration and its benchmark solves such problems
one after the other, each with a fresh clipper::CLIPPER (reference benchmarks/main.cpp:206-208, 254-270).  ``BatchCLIPPER``
takes the whole list: one CTA per problem scores the pairs, builds the compact copy and runs the solver, no device-wide
synchronisation anywhere; results per problem are those of ``CLIPPER.score_pairwise_consistency`` + ``solve``.
"""
import ctypes as C

import numpy as np

from . import _capi
from .api import EuclideanDistance, PointNormalDistance, Params, Solution, _check_f64


class BatchCLIPPER:
    """invariant: EuclideanDistance (data 2 x n or 3 x n) or PointNormalDistance (6 x n); params: Params (rounding NONZERO or DSD_HEU)"""

    def __init__(self, invariant, params=None, device=0):
        if not isinstance(invariant, (EuclideanDistance, PointNormalDistance)):
            raise TypeError("a batch scores with the built-in invariants (EuclideanDistance, PointNormalDistance)")
        self._lib = _capi.load()
        self._b = C.c_void_p()
        rc = self._lib.clp_batch_create(int(device), C.byref(self._b))
        if rc != _capi.OK:
            msg = self._lib.clp_batch_last_error(None)
            raise _capi.ClipperError(rc, msg.decode() if msg else "clp_batch_create failed")
        self._invariant = invariant
        self._params = params if params is not None else Params()

    def __del__(self):
        try:
            if getattr(self, "_b", None):
                self._lib.clp_batch_destroy(self._b)
                self._b = None
        except Exception:
            pass

    def _check(self, rc):
        if rc != _capi.OK:
            msg = self._lib.clp_batch_last_error(self._b)
            raise _capi.ClipperError(rc, msg.decode() if msg else "")

    def info(self):
        """(CTAs of the last launch, HBM scratch bytes, stored affinities i<j over all problems)"""
        a, b, c = C.c_int32(), C.c_int64(), C.c_int64()
        self._check(self._lib.clp_batch_info(self._b, C.byref(a), C.byref(b), C.byref(c)))
        return int(a.value), int(b.value), int(c.value)

    def solve_many(self, problems):
        """problems: sequence of dicts with D1 (d, n1) float64, D2 (d, n2) float64, A (m, 2) int32 or None (all-to-all)
        and u0 (m,) float64.  Returns the list of api.Solution in input order (``.kernel_ms`` = device time of the whole
        batch launch)."""
        problems = list(problems)
        n = len(problems)
        if n == 0:
            return []
        self._check(self._lib.clp_batch_set_params(self._b, C.byref(self._params._pod())))
        keep = []  # keeps the converted arrays alive during the call
        vp = C.c_void_p
        D1p, D2p, Ap, u0p, up, np_ = (vp * n)(), (vp * n)(), (vp * n)(), (vp * n)(), (vp * n)(), (vp * n)()
        n1, n2, m = (C.c_int64 * n)(), (C.c_int64 * n)(), (C.c_int64 * n)()
        us, nodes = [], []
        d = None
        for k, p in enumerate(problems):
            D1 = np.asfortranarray(_check_f64(p["D1"], "D1")); D2 = np.asfortranarray(_check_f64(p["D2"], "D2"))
            if D1.ndim != 2 or D2.ndim != 2 or D1.shape[0] != D2.shape[0]:
                raise ValueError("D1 and D2 must be (d, n1) and (d, n2)")
            if d is None:
                d = D1.shape[0]
            elif d != D1.shape[0]:
                raise ValueError("all problems of a batch must have the same data dimension")
            A = p.get("A")
            if A is None or np.size(A) == 0:
                mk = D1.shape[1] * D2.shape[1]; Ak = None
            else:
                if not isinstance(A, np.ndarray) or A.dtype != np.int32 or A.ndim != 2 or A.shape[1] != 2:
                    raise TypeError("A must be an (m, 2) numpy int32 array")
                Ak = np.asfortranarray(A); mk = Ak.shape[0]
            u0 = np.ascontiguousarray(_check_f64(p["u0"], "u0")).reshape(-1)
            if u0.size != mk:
                raise ValueError("problem %d: u0 has %d entries, expected %d" % (k, u0.size, mk))
            uk = np.zeros(mk, np.float64); nk = np.zeros(max(mk, 1), np.int32)
            keep += [D1, D2, Ak, u0]
            us.append(uk); nodes.append(nk)
            D1p[k], D2p[k], u0p[k] = D1.ctypes.data, D2.ctypes.data, u0.ctypes.data
            Ap[k] = Ak.ctypes.data if Ak is not None else None
            up[k], np_[k] = uk.ctypes.data, nk.ctypes.data
            n1[k], n2[k], m[k] = D1.shape[1], D2.shape[1], (mk if Ak is not None else 0)
        sols = (_capi.ClpSolution * n)()
        inv = self._invariant
        if isinstance(inv, EuclideanDistance):
            q = inv.params()
            rc = self._lib.clp_batch_solve_euclidean(self._b, n, int(d), D1p, n1, D2p, n2, Ap, m, u0p,
                                                     float(q.sigma), float(q.epsilon), float(q.mindist), sols, up, np_)
        else:
            if d != 6:
                raise ValueError("PointNormalDistance expects 6 x n data")
            q = inv.params()
            rc = self._lib.clp_batch_solve_pointnormal(self._b, n, D1p, n1, D2p, n2, Ap, m, u0p,
                                                       float(q.sigp), float(q.epsp), float(q.sign), float(q.epsn), sols, up, np_)
        self._check(rc)
        out = []
        for k in range(n):
            s = sols[k]
            o = Solution()
            o.t, o.ifinal, o.score = s.t, s.ifinal, s.score
            o.nodes = nodes[k][: s.n_nodes].tolist()
            o.u0, o.u = keep[4 * k + 3], us[k]
            o.d_final, o.n_evals, o.n_matvec, o.n_inner, o.kernel_ms = s.d_final, s.n_evals, s.n_matvec, s.n_inner, s.kernel_ms
            o.prof_ms = (s.prof_matvec_ms, s.prof_combine_ms, s.prof_exchange_ms)
            out.append(o)
        return out
