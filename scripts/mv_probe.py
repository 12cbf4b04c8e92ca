"""Timing probe: stand-alone compact mat-vec of config 2 (used with CLP_PROBE_NO_CONFLICT to bound the cost of
shared-memory bank conflicts; results are wrong under that flag)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import clipper_b200 as clipperpy
from clipper_b200 import _capi, datagen
m = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
prob = datagen.config_problem("c2", m); cfg = prob["cfg"]
ip = clipperpy.invariants.EuclideanDistanceParams(); ip.sigma, ip.epsilon = cfg["sigma"], cfg["epsilon"]
clip = clipperpy.CLIPPER(clipperpy.invariants.EuclideanDistance(ip), clipperpy.Params(), device=0)
clip.score_pairwise_consistency(prob["D1"], prob["D2"], prob["A"])
L = _capi.load(); h = clip.handle
nnz, pass_bytes = clip.sparse_info()
v = torch.rand(m, dtype=torch.float64, device="cuda"); y = torch.empty_like(v)
ms = C.c_double()
_capi.check(h, L.clp_matvec_dev(h, v.data_ptr(), 1.0, y.data_ptr(), None, None, 5, C.byref(ms)))
_capi.check(h, L.clp_matvec_dev(h, v.data_ptr(), 1.0, y.data_ptr(), None, None, 100, C.byref(ms)))
print("mode", clip.dense_mode(), "nnz", nnz, "pass_bytes", pass_bytes, "ms", ms.value, "GB/s", pass_bytes / ms.value / 1e6,
      "probe" if os.environ.get("CLP_PROBE_NO_CONFLICT") else "")
