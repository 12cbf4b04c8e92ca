"""The C-ABI library loads without a GPU and exports every symbol include/clipper_b200.h declares.
No compute call is made here."""
import ctypes
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "clipper_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(clp_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(built):
    from clipper_b200 import _capi
    L = ctypes.CDLL(_capi.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), "missing export: " + n
    assert sorted(_capi.SYMBOLS) == names


def test_host_utils_match_oracle(built):
    import clipper_b200 as clipperpy
    from oracle import clipper_oracle as orc
    for n in (2, 3, 12, 57):
        for k in range(n * (n - 1) // 2):
            assert clipperpy.utils.k2ij(k, n) == orc.k2ij(k, n)
    assert np.array_equal(clipperpy.utils.create_all_to_all(4, 3), orc.create_all_to_all(4, 3))
    rng = np.random.default_rng(1)
    for t in range(50):
        x = np.round(rng.random(40), 1)  # many ties
        k = int(rng.integers(0, 45))
        assert clipperpy.utils.find_indices_of_k_largest(x, k).tolist() == orc.find_k_largest(x, k).tolist()
        assert clipperpy.utils.find_indices_where_above_threshold(x, 0.5).tolist() == orc.find_above(x, 0.5).tolist()
    assert clipperpy.utils.find_indices_of_k_largest(np.array([5.0, 5, 5, 7]), 2).tolist() == [3, 1]


def test_dsd_known_answer(built):
    # reference test/dsd_test.cpp:14-80
    import clipper_b200 as clipperpy
    import fixtures as fx
    M, _ = fx.m20()
    assert clipperpy.dsd.solve(M) == fx.DSD_NODES_20
    assert clipperpy.dsd.solve(M, [0, 1, 3, 5, 7, 12, 14, 15, 19]) == fx.DSD_NODES_20


def test_create_fails_loudly_without_gpu(built):
    """No CUDA device here -> clp_create must return an error, never a CPU fallback."""
    import torch
    if torch.cuda.is_available():
        return
    import pytest
    import clipper_b200 as clipperpy
    with pytest.raises(clipperpy.ClipperError):
        clipperpy.CLIPPER(clipperpy.invariants.EuclideanDistance(clipperpy.invariants.EuclideanDistanceParams()),
                          clipperpy.Params())


def test_params_defaults_mirror_reference(built):
    # reference include/clipper/clipper.h:27-60
    import clipper_b200 as clipperpy
    p = clipperpy.Params()
    assert (p.tol_u, p.tol_F, p.tol_Fop) == (1e-8, 1e-9, 1e-10)
    assert (p.maxiniters, p.maxoliters, p.maxlsiters) == (200, 1000, 99)
    assert (p.beta, p.eps, p.affinityeps) == (0.25, 1e-9, 1e-4)
    assert p.rescale_u0 is True and p.rounding == clipperpy.Rounding.DSD_HEU
    e = clipperpy.invariants.EuclideanDistanceParams()
    assert (e.sigma, e.epsilon, e.mindist) == (0.01, 0.06, 0.0)
    q = clipperpy.invariants.PointNormalDistanceParams()
    assert (q.sigp, q.epsp, q.sign, q.epsn) == (0.5, 0.5, 0.10, 0.35)


def test_batch_create_fails_loudly_without_gpu(built):
    """the batch path has no CPU fallback either"""
    import torch
    if torch.cuda.is_available():
        return
    import pytest
    import clipper_b200 as clipperpy
    ip = clipperpy.invariants.EuclideanDistanceParams()
    with pytest.raises(clipperpy.ClipperError):
        clipperpy.BatchCLIPPER(clipperpy.invariants.EuclideanDistance(ip), clipperpy.Params())
    with pytest.raises(TypeError):
        clipperpy.BatchCLIPPER(object(), clipperpy.Params())
