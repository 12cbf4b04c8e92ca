import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import clipper_b200 as clp
from oracle import clipper_oracle as orc
for m in (127, 129):
    rng = np.random.default_rng(m)
    n = 64
    D1 = np.asfortranarray(rng.random((3, n))); D2 = np.asfortranarray(D1 + 0.001 * rng.standard_normal((3, n)))
    A = np.stack([rng.permutation(n)[:m] if m <= n else rng.integers(0, n, m),
                  rng.permutation(n)[:m] if m <= n else rng.integers(0, n, m)], axis=1).astype(np.int32)
    A[: m // 2, 1] = A[: m // 2, 0]
    for storage in (0, 1):
        o = orc.Oracle(); o.score_euclidean(D1, D2, A, sigma=0.01, epsilon=0.05)
        u0 = rng.random(m) + 0.1
        so = o.solve(u0)
        for mode in (0, 1, 2):
            ip = clp.invariants.EuclideanDistanceParams(); ip.sigma, ip.epsilon = 0.01, 0.05
            c = clp.CLIPPER(clp.invariants.EuclideanDistance(ip), clp.Params(), storage=storage)
            c.set_dense_mode(mode)
            c.score_pairwise_consistency(D1, D2, A)
            c.solve(u0); s = c.get_solution()
            v = np.random.default_rng(1).random(m)
            y, Mv, Cv = c.matvec(v, 0.5)
            print("m", m, "st", storage, "mode", mode, "F %.12f (oracle %.12f)" % (s.score, so.score), "evals", s.n_evals, so.n_evals,
                  "nodes", len(s.nodes), len(so.nodes), "same", sorted(s.nodes) == sorted(so.nodes.tolist()),
                  "mv err", np.abs(Mv - o.matvec(v, 0)).max(), np.abs(Cv - o.matvec(v, 1)).max())
