"""The fp32 screening pass of the scoring kernel (clp_kernels.cuh, FILTER instances of score_tile_kernel) may only
drop a pair when the exact fp64 consistency test |l1 - l2| < eps (reference euclidean_distance.cpp:13-31,
pointnormal_distance.cpp:13-35) must fail too.  This file re-states the kernel's fp32 arithmetic in numpy --
inputs rounded to float32, float32 subtraction, the fused sum of squares, a square root that is allowed to be off
by 2 ulp in either direction (sqrt.approx.f32), float32 subtraction of the two lengths -- and checks the
threshold  eps + 1024 * 2^-24 * R  (R = largest |coordinate|) against the exact test on data built to sit on the
decision boundary, at coordinate scales from 1e-3 to 1e6.  CPU only: it pins the error analysis the kernel
comment quotes, not the kernel itself (tests/test_gpu_parity.py compares the kernel with the oracle)."""
import numpy as np
import pytest

U = 2.0 ** -24


def fma32(a, b, c):
    """float32 fma: a*b is exact in float64 (24+24 bits), one more rounding to float64 is far below float32's ulp"""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def length32(p, q, ulp_shift):
    """|p - q| the way the kernel computes it; ulp_shift moves the square root by that many float32 ulps"""
    d = p - q                                   # float32 subtraction
    s = (d[:, 0] * d[:, 0]).astype(np.float32)  # x*x
    s = fma32(d[:, 1], d[:, 1], s)
    s = fma32(d[:, 2], d[:, 2], s)
    r = np.sqrt(s.astype(np.float64)).astype(np.float32)
    for _ in range(abs(ulp_shift)):
        r = np.nextafter(r, np.float32(np.inf if ulp_shift > 0 else -np.inf))
    return r


def threshold(eps, R):
    t = (eps + 1024.0 * U * float(np.float32(R))) * (1.0 + 2.0 ** -20)
    f = np.float32(t)
    return f if float(f) >= t else np.nextafter(f, np.float32(np.inf))  # __double2float_ru


@pytest.mark.parametrize("scale", [1e-3, 1.0, 37.5, 1e3, 1e6])
@pytest.mark.parametrize("eps_rel", [1e-7, 1e-5, 1e-3, 5e-2])
def test_no_consistent_pair_is_screened_out(scale, eps_rel):
    rng = np.random.default_rng(int(scale * 1000) % 9973 + int(eps_rel * 1e9) % 7919)
    n = 200_000
    eps = eps_rel * scale
    a1 = rng.uniform(-scale, scale, (n, 3)); b1 = rng.uniform(-scale, scale, (n, 3))
    l1 = np.linalg.norm(a1 - b1, axis=1)
    # second view: same direction, length changed by something around eps (inside, outside and on the boundary)
    delta = eps * rng.choice([0.0, 0.5, 0.999999, 1.0, 1.000001, 1.5, 3.0], n) * rng.choice([-1.0, 1.0], n)
    dirn = (a1 - b1) / np.maximum(l1, 1e-300)[:, None]
    shift = rng.uniform(-scale, scale, (n, 3)) * 0.25
    a2 = a1 + shift
    b2 = a2 - dirn * np.maximum(l1 + delta, 0.0)[:, None]
    # a few degenerate pairs: coincident points, tiny and huge lengths
    a1[:16] = b1[:16]; a2[16:32] = b2[16:32]
    R = max(np.abs(x).max() for x in (a1, b1, a2, b2))
    # exact test, as the oracle / the kernel's fp64 path compute it (sequential sum of squares, IEEE sqrt)
    def exact_len(p, q):
        d = p - q
        return np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2])
    c = np.abs(exact_len(a1, b1) - exact_len(a2, b2))
    consistent = c < eps
    assert consistent.sum() > n // 10 and (~consistent).sum() > n // 10
    f = [x.astype(np.float32) for x in (a1, b1, a2, b2)]
    thr = threshold(eps, R)
    for s1 in (-2, 0, 2):
        for s2 in (-2, 0, 2):
            c32 = np.abs(length32(f[0], f[1], s1) - length32(f[2], f[3], s2))
            dropped = c32 >= thr
            assert not np.any(dropped & consistent), (scale, eps_rel, s1, s2)
            # and the bound the kernel comment states: the screened quantity is within 90 u R of the exact one
            assert np.max(np.abs(c32.astype(np.float64) - c)) < 90.0 * U * R


def test_screening_still_rejects_what_is_clearly_inconsistent():
    """power check at the scale of BASELINE.json's config 2 (unit cloud, eps = 0.05): the margin is 6e-5"""
    rng = np.random.default_rng(5)
    n = 100_000
    a1 = rng.uniform(-1, 1, (n, 3)); b1 = rng.uniform(-1, 1, (n, 3))
    a2 = rng.uniform(-1, 1, (n, 3)); b2 = rng.uniform(-1, 1, (n, 3))
    c = np.abs(np.linalg.norm(a1 - b1, axis=1) - np.linalg.norm(a2 - b2, axis=1))
    f = [x.astype(np.float32) for x in (a1, b1, a2, b2)]
    thr = threshold(0.05, 1.0)
    assert float(thr) - 0.05 < 1e-4
    kept = np.abs(length32(f[0], f[1], 0) - length32(f[2], f[3], 0)) < thr
    assert not np.any(kept & (c > 0.0502))
    assert np.all(kept[c < 0.05])


def test_non_finite_input_is_never_screened_out():
    """NaN / Inf make the kernel's comparison `|l1 - l2| >= thr` false (the pair goes to the exact path); an
    infinite or NaN scale makes the threshold infinite or NaN with the same effect"""
    for R in (np.inf, np.nan):
        with np.errstate(invalid="ignore", over="ignore"):
            t = np.float32((0.05 + 1024.0 * U * R) * (1.0 + 2.0 ** -20))
            for c32 in (np.float32(0.0), np.float32(1e30), np.float32(np.nan)):
                assert not (c32 >= t) or (np.isinf(t) and np.isinf(c32))
    with np.errstate(invalid="ignore"):
        assert not (np.float32(np.nan) >= np.float32(0.05))
