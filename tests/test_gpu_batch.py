"""Batches of small problems in one launch (clp_batch_*): every problem against the ORACLE
(identical inlier set, objective 1e-5) and against the single-problem path on the same inputs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def clp(built):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import clipper_b200 as clipperpy
    return clipperpy


def _euclid(clp, sigma, epsilon, mindist=0.0):
    ip = clp.invariants.EuclideanDistanceParams(); ip.sigma, ip.epsilon, ip.mindist = sigma, epsilon, mindist
    return clp.invariants.EuclideanDistance(ip)


def _problems(sizes, seed0, rho=0.9):
    """seeded problems on the benchmark cloud; only the points the associations touch travel (keeps inputs small)"""
    from clipper_b200 import datagen
    out = []
    for k, m in enumerate(sizes):
        prob = datagen.euclidean_problem(m, rho, seed0 + k)
        A = prob["A"]
        used1, inv1 = np.unique(A[:, 0], return_inverse=True)
        used2, inv2 = np.unique(A[:, 1], return_inverse=True)
        out.append(dict(D1=np.asfortranarray(prob["D1"][:, used1]), D2=np.asfortranarray(prob["D2"][:, used2]),
                        A=np.asfortranarray(np.stack([inv1, inv2], axis=1).astype(np.int32)), u0=prob["u0"]))
    return out


def test_batch_matches_oracle_and_single(clp):
    from oracle import clipper_oracle as orc
    sizes = [64, 100, 256, 333, 512, 777, 1000, 1024, 1500, 2048, 65, 129, 12, 4, 640, 900] * 3
    probs = _problems(sizes, 500)
    sigma, eps = 0.015, 0.05
    b = clp.BatchCLIPPER(_euclid(clp, sigma, eps), clp.Params())
    sols = b.solve_many(probs)
    assert len(sols) == len(probs)
    ctas, scratch, nnz = b.info()
    assert ctas >= 1 and scratch > 0
    nnz_or = 0
    for k, (p, s) in enumerate(zip(probs, sols)):
        o = orc.Oracle(); o.score_euclidean(p["D1"], p["D2"], p["A"], sigma=sigma, epsilon=eps)
        so = o.solve(p["u0"])
        nnz_or += o.nnz(0)
        assert sorted(s.nodes) == sorted(so.nodes.tolist()), "problem %d (m=%d): inlier set differs" % (k, sizes[k])
        assert abs(s.score - so.score) <= 1e-5 * max(1.0, abs(so.score)), (k, s.score, so.score)
        assert np.abs(s.u - so.u).max() <= 1e-4
        assert s.n_matvec == s.n_evals + 2
    assert nnz == nnz_or  # the batch stored exactly the oracle's affinities
    # the single-problem path on a few of them: same algorithm, different CTA count -> same decisions
    for k in (0, 6, 9, 12):
        c = clp.CLIPPER(_euclid(clp, sigma, eps), clp.Params())
        c.score_pairwise_consistency(probs[k]["D1"], probs[k]["D2"], probs[k]["A"])
        c.solve(probs[k]["u0"]); s1 = c.get_solution()
        assert sorted(s1.nodes) == sorted(sols[k].nodes) and abs(s1.score - sols[k].score) <= 1e-9 * max(1.0, abs(s1.score))
        assert s1.n_evals == sols[k].n_evals and s1.ifinal == sols[k].ifinal
    # reproducible bit for bit from call to call (problems are drawn dynamically, slots differ)
    again = b.solve_many(probs)
    assert all(a.u.tobytes() == s.u.tobytes() and a.score == s.score and a.nodes == s.nodes for a, s in zip(again, sols))


def test_batch_all_to_all_and_rounding(clp):
    from oracle import clipper_oracle as orc
    import fixtures as fx
    model, data = fx.toy_problem()   # reference test/clipper_test.cpp:34-66, all-to-all hypothesis
    p = clp.Params(); p.rounding = clp.Rounding.NONZERO
    b = clp.BatchCLIPPER(_euclid(clp, 0.01, 0.06), p)
    probs = [dict(D1=model, D2=data, A=None, u0=np.full(12, 0.5)) for _ in range(5)]
    sols = b.solve_many(probs)
    o = orc.Oracle(orc.default_params(rounding=orc.NONZERO)); o.score_euclidean(model, data, None)
    so = o.solve(np.full(12, 0.5))
    for s in sols:
        assert s.nodes == so.nodes.tolist() and abs(s.score - so.score) <= 1e-9 * abs(so.score)
    # DSD is refused loudly, never silently replaced
    pd = clp.Params(); pd.rounding = clp.Rounding.DSD
    with pytest.raises(clp.ClipperError):
        clp.BatchCLIPPER(_euclid(clp, 0.01, 0.06), pd).solve_many(probs)
    # too large a problem is refused
    from clipper_b200 import datagen
    big = datagen.euclidean_problem(5000, 0.9, 3)
    with pytest.raises(clp.ClipperError):
        b.solve_many([dict(D1=big["D1"], D2=big["D2"], A=big["A"], u0=big["u0"])])


def test_batch_pointnormal(clp):
    from oracle import clipper_oracle as orc
    from clipper_b200 import datagen
    ip = clp.invariants.PointNormalDistanceParams()
    probs = []
    for k, m in enumerate((200, 400, 800)):
        pr = datagen.pointnormal_problem(m, 0.9, 700 + k)
        probs.append(dict(D1=pr["D1"], D2=pr["D2"], A=pr["A"], u0=pr["u0"]))
    sols = clp.BatchCLIPPER(clp.invariants.PointNormalDistance(ip), clp.Params()).solve_many(probs)
    for p, s in zip(probs, sols):
        o = orc.Oracle(); o.score_pointnormal(p["D1"], p["D2"], p["A"], sigp=ip.sigp, epsp=ip.epsp, sign=ip.sign, epsn=ip.epsn)
        so = o.solve(p["u0"])
        assert sorted(s.nodes) == sorted(so.nodes.tolist()) and abs(s.score - so.score) <= 1e-5 * abs(so.score)
