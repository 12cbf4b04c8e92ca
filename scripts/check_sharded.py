#!/usr/bin/env python
"""torchrun --nproc-per-node N scripts/check_sharded.py [m ...]: the row-sharded solve (one process per GPU, NCCL
plumbing, in-kernel NVLink exchange) against the CPU oracle: identical inlier set, |dF|/F <= 1e-5, all ranks
bit-identical.  One JSON line per problem on rank 0 (appended to gpurun_out/sharded_vs_oracle.jsonl)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import clipper_b200 as clp
    from clipper_b200 import datagen, distributed as cd
    sizes = [int(x) for x in sys.argv[1:]] or [3000, 20000]
    for m in sizes:
        name = "c4" if m > 27648 else "c2"
        prob = datagen.config_problem(name, m); cfg = prob["cfg"]
        ip = clp.invariants.EuclideanDistanceParams(); ip.sigma, ip.epsilon = cfg["sigma"], cfg["epsilon"]
        c = cd.ShardedCLIPPER(clp.invariants.EuclideanDistance(ip), clp.Params())
        c.score_pairwise_consistency(prob["D1"], prob["D2"], prob["A"])
        for rep in range(2):
            c.solve(prob["u0"])
        s = c.get_solution()
        sel = c.get_selected_associations()
        # all ranks bit-identical
        dev = torch.device("cuda", local)
        h = torch.tensor([float(np.frombuffer(s.u.tobytes(), dtype=np.uint8).astype(np.float64).sum()), s.score, float(len(s.nodes))],
                         dtype=torch.float64, device=dev)
        lo, hi = h.clone(), h.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same_ranks = bool(torch.equal(lo, hi))
        if rank == 0:
            from oracle import clipper_oracle as orc
            rec = dict(m=m, world=world, mode=c.dense_mode(), evals=int(s.n_evals), kernel_ms=s.kernel_ms, ranks_identical=same_ranks,
                       F=s.score, n_nodes=len(s.nodes))
            if m <= 30000:
                o = orc.Oracle(); o.score_euclidean(prob["D1"], prob["D2"], prob["A"], sigma=cfg["sigma"], epsilon=cfg["epsilon"])
                so = o.solve(prob["u0"])
                rec.update(same_inlier_set=sorted(s.nodes) == sorted(so.nodes.tolist()), rel_dF=abs(s.score - so.score) / abs(so.score),
                           evals_oracle=int(so.n_evals), max_du=float(np.abs(s.u - so.u).max()),
                           selected_ok=bool(np.array_equal(np.sort(sel, axis=0), np.sort(o.get_initial_associations()[so.nodes], axis=0))))
                assert rec["same_inlier_set"] and rec["rel_dF"] <= 1e-5 and rec["selected_ok"], rec
            assert same_ranks, rec
            print(json.dumps(rec), flush=True)
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", "sharded_vs_oracle.jsonl"), "a") as f:
                f.write(json.dumps(rec) + "\n")
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
