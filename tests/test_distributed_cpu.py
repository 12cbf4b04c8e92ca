"""Host-side logic of the row-block-sharded path, world_size 2 (and 3) on CPU with gloo:
partitioning, blob all-gather plumbing, and the collective formulation of the exchange step
(all-gather of disjoint slices == all-reduce of zero-padded vectors)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, m, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from clipper_b200 import distributed as cd, _capi
        r0, n = cd.shard_rows(m, rank, world)
        # (1) every rank derives the same global partition
        rows = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(rows, torch.tensor([r0, n], dtype=torch.int64))
        nxt = 0
        for t in rows:
            assert int(t[0]) == nxt; nxt += int(t[1])
        assert nxt == m
        # (2) blob plumbing: fixed-size opaque blobs come back in rank order
        nb = int(_capi.load().clp_shard_blob_bytes())
        mine = torch.full((nb,), rank + 1, dtype=torch.uint8)
        allb = torch.empty(world * nb, dtype=torch.uint8)
        dist.all_gather_into_tensor(allb, mine)
        for r in range(world):
            assert (allb[r * nb:(r + 1) * nb] == r + 1).all()
        # (3) the exchange step as a collective: every rank owns y[r0:r0+n] of a seeded mat-vec
        rng = np.random.default_rng(7)
        Mfull = rng.random((m, m)); Mfull = Mfull + Mfull.T
        v = rng.random(m)
        y_loc = torch.from_numpy(Mfull[r0:r0 + n] @ v)
        y_gather = cd.gathered_matvec(y_loc)
        y_pad = torch.zeros(m, dtype=torch.float64); y_pad[r0:r0 + n] = y_loc
        dist.all_reduce(y_pad)
        assert torch.equal(y_gather, y_pad)  # exactly one non-zero contributor per entry -> bit exact
        assert np.allclose(y_gather.numpy(), Mfull @ v, rtol=1e-13)
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,m", [(2, 1000), (2, 33), (3, 130)])
def test_gloo_sharded_host_logic(built, world, m):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, m, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in ps:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def test_partition_exact_for_many_sizes(built):
    from clipper_b200 import distributed as cd
    for m in (1, 2, 31, 32, 33, 64, 1000, 20000, 80000, 131072):
        for world in (1, 2, 3, 4, 8):
            assert cd.partition_is_exact(m, world), (m, world)
            for r in range(world):
                r0, n = cd.shard_rows(m, r, world)
                assert r0 % 32 == 0  # shards never split a 32-row tile


def test_interleaved_shards_same_answer_and_balanced(built):
    """opt-in byte balancing of the row shards: dealing the associations round-robin is a bijection that keeps the
    shard sizes, spreads a block of dense rows evenly, and -- solved by the oracle -- gives the same inlier set
    once mapped back"""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden
    from clipper_b200 import distributed as cd
    from oracle.clipper_oracle import Oracle
    for m, world in ((1000, 8), (333, 4), (20000, 8), (64, 2)):
        p = cd.interleave_permutation(m, world)
        assert sorted(p.tolist()) == list(range(m))
        last = np.arange(m - m // 20, m)  # the generator's inlier block: the last 5 % of the associations
        for r in range(world):
            r0, n = cd.shard_rows(m, r, world)
            got = int(((p[last] >= r0) & (p[last] < r0 + n)).sum())
            assert abs(got - len(last) * n / m) <= 1.0 + 1e-9, (m, world, r, got)
    g = make_golden.load("euclid_m300")
    m = g["A"].shape[0]
    p = cd.interleave_permutation(m, 4)
    Ap, u0p = cd.permute_problem(g["A"], g["u0"], p)
    o = Oracle(); o.score_euclidean(g["D1"], g["D2"], Ap, **g["params"])
    s = o.solve(u0p)
    nodes, u = cd.unpermute_solution(s.nodes, s.u, p)
    assert sorted(nodes.tolist()) == sorted(g["nodes"].tolist())
    assert abs(s.score - float(g["score"])) <= 1e-12 * float(g["score"])
    assert np.allclose(u, g["u"], rtol=0, atol=1e-12)
