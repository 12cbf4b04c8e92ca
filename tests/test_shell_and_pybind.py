"""The reference's C++ API (include/clipper/*.h over the C-ABI) and its pybind11 module `clipperpy`.
CPU part: the shell links, the GPU-free C++ tests pass, the module imports with the reference's names.
GPU part: the reference's C++ tests restated in tests/cpp/shell_tests.cpp, and the example
notebook's call sequence (reference examples/python/ex4_bunny.ipynb) through `clipperpy`."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "clipper_b200", "lib")


def _clipperpy():
    if LIB not in sys.path:
        sys.path.insert(0, LIB)
    import clipperpy
    return clipperpy


def test_cpp_shell_cpu_only(built):
    out = subprocess.run([os.path.join(LIB, "shell_tests"), "--cpu-only"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "0 failures" in out.stdout


def test_clipperpy_surface(built):
    # reference bindings/python/py_clipper.cpp:116-232
    cp = _clipperpy()
    assert cp.__version__
    for name in ("Invariant", "PairwiseInvariant", "EuclideanDistanceParams", "EuclideanDistance",
                 "PointNormalDistanceParams", "PointNormalDistance"):
        assert hasattr(cp.invariants, name)
    for name in ("create_all_to_all", "k2ij"):
        assert hasattr(cp.utils, name) and hasattr(cp.dsd, name)
    for name in ("MCParams", "SDPParams", "Rounding", "Params", "Solution", "CLIPPER"):
        assert hasattr(cp, name)
    for meth in ("score_pairwise_consistency", "solve", "solve_as_maximum_clique", "solve_as_msrc_sdr",
                 "get_initial_associations", "get_selected_associations", "get_solution", "get_affinity_matrix",
                 "get_constraint_matrix", "set_matrix_data", "set_parallelize"):
        assert hasattr(cp.CLIPPER, meth)
    p = cp.Params()
    assert (p.tol_u, p.maxiniters, p.beta, p.rounding) == (1e-8, 200, 0.25, cp.Rounding.DSD_HEU)
    ip = cp.invariants.EuclideanDistanceParams()
    assert "sigma=0.01" in repr(ip)
    inv = cp.invariants.EuclideanDistance(ip)
    a, b = np.zeros(3), np.array([1.0, 0, 0])
    assert inv(a, b, a, b) == 1.0
    assert cp.utils.k2ij(0, 4) == (0, 1)
    assert cp.utils.create_all_to_all(2, 3).shape == (6, 2)
    import fixtures as fx
    M, _ = fx.m20()
    assert cp.dsd.solve(M) == fx.DSD_NODES_20


@pytest.mark.gpu
def test_cpp_shell_reference_tests(built):
    out = subprocess.run([os.path.join(LIB, "shell_tests")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "9 tests" in out.stdout and "0 failures" in out.stdout


@pytest.mark.gpu
def test_clipperpy_notebook_flow(built):
    """reference examples/python/ex4_bunny.ipynb cells 3-6 and 13-15 (built-in and Python-subclassed invariant)"""
    cp = _clipperpy()
    from clipper_b200 import datagen
    prob = datagen.config_problem("c1")
    cfg = prob["cfg"]
    iparams = cp.invariants.EuclideanDistanceParams()
    iparams.sigma, iparams.epsilon = cfg["sigma"], cfg["epsilon"]
    invariant = cp.invariants.EuclideanDistance(iparams)
    params = cp.Params()
    params.rounding = cp.Rounding.DSD_HEU
    clipper = cp.CLIPPER(invariant, params)
    clipper.score_pairwise_consistency(prob["D1"], prob["D2"], prob["A"])
    clipper.solve()
    Ain = clipper.get_selected_associations()
    assert Ain.shape[1] == 2 and Ain.shape[0] > 10
    assert (Ain[:, 0] == Ain[:, 1]).mean() > 0.95  # inliers are (p, p)
    sol = clipper.get_solution()
    assert sol.u.shape == (cfg["m"],) and abs(np.linalg.norm(sol.u) - 1) < 1e-12
    M = clipper.get_affinity_matrix()
    assert M.shape == (cfg["m"], cfg["m"]) and np.array_equal(M, M.T)
    with pytest.raises(TypeError):
        clipper.score_pairwise_consistency(prob["D1"].astype(np.float32), prob["D2"], prob["A"])

    class Custom(cp.invariants.PairwiseInvariant):
        def __init__(self, sigma, epsilon):
            cp.invariants.PairwiseInvariant.__init__(self)
            self.sigma, self.epsilon = sigma, epsilon

        def __call__(self, ai, aj, bi, bj):
            c = abs(np.linalg.norm(ai - aj) - np.linalg.norm(bi - bj))
            return float(np.exp(-0.5 * c * c / self.sigma ** 2)) if c < self.epsilon else 0.0

    sub = slice(cfg["m"] - 60, cfg["m"])  # 60 associations keep the per-pair Python loop short
    A = np.asfortranarray(prob["A"][sub])
    c2 = cp.CLIPPER(Custom(cfg["sigma"], cfg["epsilon"]), params)
    c2.score_pairwise_consistency(prob["D1"], prob["D2"], A)
    c3 = cp.CLIPPER(invariant, params)
    c3.score_pairwise_consistency(prob["D1"], prob["D2"], A)
    assert np.allclose(c2.get_affinity_matrix(), c3.get_affinity_matrix(), rtol=0, atol=1e-7)
    c2.solve(np.ones(60)); c3.solve(np.ones(60))
    assert c2.get_solution().nodes == c3.get_solution().nodes
