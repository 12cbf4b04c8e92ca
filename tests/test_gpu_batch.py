"""Batched small problems: clp_set_grid_cap + clipper_b200.batch.BatchSolver.
OPT-IN (CLP_TEST_EXPERIMENTAL=1): written after round 1's GPU budget was spent, not yet run on hardware."""
import os

import numpy as np
import pytest

pytestmark = [
    pytest.mark.gpu,
    pytest.mark.skipif(os.environ.get("CLP_TEST_EXPERIMENTAL") != "1",
                       reason="experimental: set CLP_TEST_EXPERIMENTAL=1 (not validated on hardware in round 1)"),
]


def _problem(m, seed):
    from clipper_b200 import datagen
    return datagen.euclidean_problem(m, 0.9, seed)


def _clipper(cap=0, mode=None):
    import clipper_b200 as clipperpy
    ip = clipperpy.invariants.EuclideanDistanceParams(); ip.sigma, ip.epsilon = 0.01, 0.02
    c = clipperpy.CLIPPER(clipperpy.invariants.EuclideanDistance(ip), clipperpy.Params())
    if mode is not None:
        c.set_dense_mode(mode)
    c.set_grid_cap(cap)
    return c


@pytest.mark.parametrize("mode", [0, 3])
@pytest.mark.parametrize("cap", [1, 3, 8, 24, 100])
def test_capped_grid_gives_the_same_answer(built, cap, mode):
    p = _problem(1000, 11)
    ref = _clipper(0, mode); ref.score_pairwise_consistency(p["D1"], p["D2"], p["A"]); ref.solve(p["u0"])
    r = ref.get_solution()
    c = _clipper(cap, mode); c.score_pairwise_consistency(p["D1"], p["D2"], p["A"]); c.solve(p["u0"])
    s = c.get_solution()
    assert s.nodes == r.nodes and s.ifinal == r.ifinal and s.n_evals == r.n_evals
    assert abs(s.score - r.score) <= 1e-9 * abs(r.score)
    assert np.allclose(s.u, r.u, rtol=0, atol=1e-10)


def test_batch_equals_one_by_one(built):
    import clipper_b200 as clipperpy
    from clipper_b200.batch import BatchSolver
    probs = [_problem(300 + 53 * k, 100 + k) for k in range(12)]
    one = _clipper(0)
    want = []
    for p in probs:
        one.score_pairwise_consistency(p["D1"], p["D2"], p["A"]); one.solve(p["u0"])
        want.append(one.get_solution())

    def inv():
        ip = clipperpy.invariants.EuclideanDistanceParams(); ip.sigma, ip.epsilon = 0.01, 0.02
        return clipperpy.invariants.EuclideanDistance(ip)

    got = BatchSolver(inv, clipperpy.Params(), workers=4, grid_cap=16).solve_many(probs)
    for g, w, p in zip(got, want, probs):
        assert g.nodes == w.nodes
        assert abs(g.score - w.score) <= 1e-9 * max(abs(w.score), 1.0)
        assert g.associations.shape == (len(w.nodes), 2)


@pytest.mark.parametrize("depth", [2, 3, 4, 6])
def test_ring_sweep_matches_the_default_sweep(built, depth, monkeypatch):
    """solver_kernel<float,5> (cp.async ring, CLP_SPARSE_RING=<depth>, read when the handle is created) against the
    default compact sweep: same decisions, same inlier set, objective to rounding"""
    p = _problem(3000, 21)
    ref = _clipper(0, 3); ref.score_pairwise_consistency(p["D1"], p["D2"], p["A"]); ref.solve(p["u0"])
    r = ref.get_solution()
    monkeypatch.setenv("CLP_SPARSE_RING", str(depth))
    c = _clipper(0, 3)
    monkeypatch.delenv("CLP_SPARSE_RING")
    c.score_pairwise_consistency(p["D1"], p["D2"], p["A"]); c.solve(p["u0"])
    s = c.get_solution()
    assert s.nodes == r.nodes and s.ifinal == r.ifinal and s.n_evals == r.n_evals
    assert abs(s.score - r.score) <= 1e-10 * abs(r.score)
    assert np.allclose(s.u, r.u, rtol=0, atol=1e-11)
