"""problems/s for a batch of small registrations: BatchSolver vs one object, one at a time.
EXPERIMENTAL, not part of bench.py's contract.  usage: python scripts/bench_batch.py [m] [count] [workers] [grid_cap]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clipper_b200 as clipperpy
from clipper_b200 import datagen
from clipper_b200.batch import BatchSolver

m = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
count = int(sys.argv[2]) if len(sys.argv) > 2 else 256
workers = int(sys.argv[3]) if len(sys.argv) > 3 else 18
cap = int(sys.argv[4]) if len(sys.argv) > 4 else 24
probs = [datagen.euclidean_problem(m, 0.9, 1000 + k) for k in range(count)]


def inv():
    ip = clipperpy.invariants.EuclideanDistanceParams(); ip.sigma, ip.epsilon = 0.01, 0.02
    return clipperpy.invariants.EuclideanDistance(ip)


one = clipperpy.CLIPPER(inv(), clipperpy.Params())
for p in probs[:4]:
    one.score_pairwise_consistency(p["D1"], p["D2"], p["A"]); one.solve(p["u0"])
t0 = time.perf_counter()
ref = []
for p in probs:
    one.score_pairwise_consistency(p["D1"], p["D2"], p["A"]); one.solve(p["u0"]); ref.append(one.get_solution())
t_one = time.perf_counter() - t0

pool = BatchSolver(inv, clipperpy.Params(), workers=workers, grid_cap=cap)
pool.solve_many(probs[: 2 * workers])
t0 = time.perf_counter()
got = pool.solve_many(probs)
t_pool = time.perf_counter() - t0
same = all(g.nodes == r.nodes for g, r in zip(got, ref))
print(json.dumps({"m": m, "problems": count, "workers": workers, "grid_cap": cap, "one_by_one_problems_per_s": count / t_one,
                  "batched_problems_per_s": count / t_pool, "speedup": t_one / t_pool, "same_inlier_sets": same}))
