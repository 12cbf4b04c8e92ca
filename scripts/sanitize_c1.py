#!/usr/bin/env python
"""compute-sanitizer target: BASELINE config 1 (m = 1000) through the single-problem path (sweep mode given as argv[1],
default auto = resident kernel) and through the one-launch batch path; prints the results so that the log shows the
run was real.  Usage: compute-sanitizer --tool memcheck|racecheck python scripts/sanitize_c1.py [mode]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import clipper_b200 as clp  # noqa: E402
from clipper_b200 import datagen  # noqa: E402

prob = datagen.config_problem("c1"); cfg = prob["cfg"]
ip = clp.invariants.EuclideanDistanceParams(); ip.sigma, ip.epsilon = cfg["sigma"], cfg["epsilon"]
c = clp.CLIPPER(clp.invariants.EuclideanDistance(ip), clp.Params())
if len(sys.argv) > 1:
    c.set_dense_mode(int(sys.argv[1]))
c.score_pairwise_consistency(prob["D1"], prob["D2"], prob["A"]); c.solve(prob["u0"])
s = c.get_solution()
print("c1 sweep mode", c.dense_mode(), "F", s.score, "nodes", len(s.nodes), "evals", s.n_evals)
b = clp.BatchCLIPPER(clp.invariants.EuclideanDistance(ip), clp.Params())
sols = b.solve_many([dict(D1=prob["D1"], D2=prob["D2"], A=prob["A"], u0=prob["u0"])] * 3)
print("batch F", [x.score for x in sols])
