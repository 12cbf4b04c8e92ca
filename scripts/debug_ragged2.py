import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import clipper_b200 as clp
from oracle import clipper_oracle as orc
m = 127
rng = np.random.default_rng(m)
n = 64
D1 = np.asfortranarray(rng.random((3, n))); D2 = np.asfortranarray(D1 + 0.001 * rng.standard_normal((3, n)))
A = np.stack([rng.integers(0, n, m), rng.integers(0, n, m)], axis=1).astype(np.int32)
A[: m // 2, 1] = A[: m // 2, 0]
u0 = rng.random(m) + 0.1
for kw in (dict(maxoliters=0), dict(maxoliters=1, maxiniters=0), dict(maxoliters=1, maxiniters=1, maxlsiters=1),
           dict(maxoliters=1, maxiniters=3, maxlsiters=2), dict(maxoliters=2, maxiniters=5)):
    p = orc.default_params(**kw); o = orc.Oracle(p); o.score_euclidean(D1, D2, A, sigma=0.01, epsilon=0.05)
    so = o.solve(u0)
    for mode in (1, 2):
        ip = clp.invariants.EuclideanDistanceParams(); ip.sigma, ip.epsilon = 0.01, 0.05
        P = clp.Params()
        for k, v in kw.items(): setattr(P, k, v)
        c = clp.CLIPPER(clp.invariants.EuclideanDistance(ip), P, storage=1)
        c.set_dense_mode(mode)
        c.score_pairwise_consistency(D1, D2, A)
        c.solve(u0); s = c.get_solution()
        print(kw, "mode", mode, "F %.15g (%.15g)" % (s.score, so.score), "d %.15g (%.15g)" % (s.d_final, so.d_final),
              "du", np.abs(s.u - so.u).max(), "evals", s.n_evals, so.n_evals)
