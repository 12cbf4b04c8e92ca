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
o = orc.Oracle(); o.score_euclidean(D1, D2, A, sigma=0.01, epsilon=0.05)
for eps in (0.0, 1e-15, 1e-13, 1e-10):
    so = o.solve(u0 * (1 + eps * np.arange(m)))
    print("oracle pert", eps, "F", so.score, "nodes", len(so.nodes), "evals", so.n_evals)
for mode in (1, 2):
    for eps in (0.0, 0.0, 1e-15, 1e-13, 1e-10):
        ip = clp.invariants.EuclideanDistanceParams(); ip.sigma, ip.epsilon = 0.01, 0.05
        c = clp.CLIPPER(clp.invariants.EuclideanDistance(ip), clp.Params(), storage=1)
        c.set_dense_mode(mode)
        c.score_pairwise_consistency(D1, D2, A)
        c.solve(u0 * (1 + eps * np.arange(m))); s = c.get_solution()
        print("mode", mode, "pert", eps, "F", s.score, "nodes", len(s.nodes), "evals", s.n_evals, "ifinal", s.ifinal)
# is the 35-clique real?  check with the oracle's matrices
ip = clp.invariants.EuclideanDistanceParams(); ip.sigma, ip.epsilon = 0.01, 0.05
c = clp.CLIPPER(clp.invariants.EuclideanDistance(ip), clp.Params(), storage=1); c.set_dense_mode(2)
c.score_pairwise_consistency(D1, D2, A); c.solve(u0); s = c.get_solution()
M = o.get_affinity_matrix(); C = o.get_constraint_matrix()
idx = np.asarray(s.nodes)
sub = C[np.ix_(idx, idx)]
print("mode2 cluster size", len(idx), "is clique in C:", bool((sub == 1).all()), "min affinity", M[np.ix_(idx, idx)].min())
idx = so.nodes
print("oracle cluster size", len(idx), "is clique:", bool((C[np.ix_(idx, idx)] == 1).all()))
