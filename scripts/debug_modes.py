import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import clipper_b200 as clp
from clipper_b200 import datagen
for m in [int(x) for x in sys.argv[1:]] or [127]:
    rng = np.random.default_rng(m)
    n = 64
    D1 = np.asfortranarray(rng.random((3, n))); D2 = np.asfortranarray(D1 + 0.001 * rng.standard_normal((3, n)))
    A = np.stack([rng.integers(0, n, m), rng.integers(0, n, m)], axis=1).astype(np.int32)
    v = rng.random(m)
    outs = []
    for mode in (0, 1, 2):
        ip = clp.invariants.EuclideanDistanceParams(); ip.sigma, ip.epsilon = 0.01, 0.05
        c = clp.CLIPPER(clp.invariants.EuclideanDistance(ip), clp.Params(), storage=1)
        c.set_dense_mode(mode)
        c.score_pairwise_consistency(D1, D2, A)
        y, Mv, Cv = c.matvec(v, 0.5)
        M = c.get_affinity_matrix() - np.eye(m); Cm = c.get_constraint_matrix() - np.eye(m)
        outs.append((Mv, Cv))
        print("m", m, "mode", mode, "max|Mv - M v|", np.abs(Mv - M @ v).max(), "max|Cv - C v|", np.abs(Cv - Cm @ v).max(),
              "bad rows", np.nonzero(np.abs(Cv - Cm @ v) > 1e-9)[0][:20])

print("---- solver")
for m in [127, 100]:
    prob = datagen.config_problem("c2", m); cfg = prob["cfg"]
    for mode in (0, 1, 2):
        ip = clp.invariants.EuclideanDistanceParams(); ip.sigma, ip.epsilon = cfg["sigma"], cfg["epsilon"]
        c = clp.CLIPPER(clp.invariants.EuclideanDistance(ip), clp.Params(), storage=1)
        c.set_dense_mode(mode)
        c.score_pairwise_consistency(prob["D1"], prob["D2"], prob["A"])
        c.solve(prob["u0"]); s = c.get_solution()
        print("m", m, "mode", mode, "F", s.score, "evals", s.n_evals, "ifinal", s.ifinal, "nodes", len(s.nodes), s.nodes[:6])
