#!/usr/bin/env python
"""BASELINE config 4 (m = 80 000) once on one GPU: score, three solves, stand-alone compact mat-vec -- the target of the
ncu capture of the segmented solver (scripts/gpu_c4_profile.sh)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import clipper_b200 as clp  # noqa: E402
from clipper_b200 import _capi, datagen  # noqa: E402

m = int(sys.argv[1]) if len(sys.argv) > 1 else None
prob = datagen.config_problem("c4", m); cfg = prob["cfg"]; m = cfg["m"]
ip = clp.invariants.EuclideanDistanceParams(); ip.sigma, ip.epsilon = cfg["sigma"], cfg["epsilon"]
c = clp.CLIPPER(clp.invariants.EuclideanDistance(ip), clp.Params())
c.score_pairwise_consistency(prob["D1"], prob["D2"], prob["A"])
for _ in range(2):
    c.solve(prob["u0"])
s = c.get_solution()
kept, pb = c.sparse_info()
L = _capi.load(); dev = torch.device("cuda:0")
v = torch.rand(m, dtype=torch.float64, device=dev); y = torch.empty_like(v); ms = C.c_double()
_capi.check(c.handle, L.clp_matvec_dev(c.handle, v.data_ptr(), 1.0, y.data_ptr(), None, None, 3, C.byref(ms)))
_capi.check(c.handle, L.clp_matvec_dev(c.handle, v.data_ptr(), 1.0, y.data_ptr(), None, None, 20, C.byref(ms)))
print("c4 m=%d mode %d: solver %.2f ms, %d evals, %.3f ms per pass = %.0f GB/s (%.3f of 6574); phases %s; stand-alone compact pass %.3f ms = %.0f GB/s (%.3f)"
      % (m, c.dense_mode(), s.kernel_ms, s.n_evals, s.kernel_ms / s.n_matvec, pb * s.n_matvec / s.kernel_ms / 1e6,
         pb * s.n_matvec / s.kernel_ms / 1e6 / 6574.1, [round(x, 2) for x in s.prof_ms], ms.value, pb / ms.value / 1e6, pb / ms.value / 1e6 / 6574.1))
