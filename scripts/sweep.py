#!/usr/bin/env python
"""BASELINE.json configs 1, 3 and 5 on one B200 (config 2 is bench.py's default, config 4 its N>1 extra):
  * c5: Md.v mat-vec sweep m = 2048 ... 131072 (fp32 storage and, up to 65536, fp64 storage): ms per pass,
        algorithmic GB/s (4 m^2 + 16 m bytes, resp. 8 m^2), fraction of the measured HBM peak.  Points whose
        matrix fits the 126 MB L2 are reported but flagged (they are not HBM measurements).
  * c1 (m=1000, Bunny-like, rho=.9) and c3 (PointNormalDistance m=10000): score + solve timings, density,
        evaluation counts, inlier precision/recall, and the CPU oracle beside them.
Writes one JSON object per line to gpurun_out/sweep_r02.jsonl and prints them."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import clipper_b200 as clp  # noqa: E402
from clipper_b200 import _capi, datagen  # noqa: E402
from bench import measured_peaks, oracle_step, cpu_cores  # noqa: E402

out_path = os.path.join(ROOT, "gpurun_out", "sweep_r02.jsonl")
os.makedirs(os.path.dirname(out_path), exist_ok=True)
fout = open(out_path, "w")
peak, _ = measured_peaks()
L = _capi.load()


def emit(d):
    s = json.dumps(d)
    print(s, flush=True)
    fout.write(s + "\n"); fout.flush()


def matvec_sweep():
    dev = torch.device("cuda:0")
    for storage, esz, sizes in ((0, 4, [2048, 4096, 8192, 16384, 20000, 32768, 65536, 131072]),
                                (1, 8, [2048, 8192, 20000, 32768, 65536])):
        for m in sizes:
            prob = datagen.config_problem("c2", m); cfg = prob["cfg"]
            ip = clp.invariants.EuclideanDistanceParams(); ip.sigma, ip.epsilon = cfg["sigma"], cfg["epsilon"]
            c = clp.CLIPPER(clp.invariants.EuclideanDistance(ip), clp.Params(), storage=storage)
            c.set_dense_mode(0)  # config 5 is the DENSE Md.v sweep: full matrix, 4 m^2 (8 m^2) bytes per pass
            t0 = time.perf_counter()
            c.score_pairwise_consistency(prob["D1"], prob["D2"], prob["A"])
            t_score = time.perf_counter() - t0
            v = torch.rand(m, dtype=torch.float64, device=dev); y = torch.empty_like(v)
            ms = C.c_double()
            _capi.check(c.handle, L.clp_matvec_dev(c.handle, v.data_ptr(), 1.0, y.data_ptr(), None, None, 5, C.byref(ms)))
            _capi.check(c.handle, L.clp_matvec_dev(c.handle, v.data_ptr(), 1.0, y.data_ptr(), None, None, 50, C.byref(ms)))
            byts = esz * m * m + 16 * m
            gbs = byts / (ms.value * 1e-3) / 1e9
            rec = {"config": "c5", "storage": "f32" if storage == 0 else "f64", "m": m, "ms_per_matvec": ms.value,
                   "algorithmic_GB": byts / 1e9, "GBps": gbs, "frac_of_measured_hbm_peak": gbs / peak,
                   "l2_resident": bool(esz * m * m < 126e6), "t_score_call_s": t_score}
            # the compact copy the solver actually sweeps (auto: resident full rows for m <= 27648, else column segments)
            c.set_dense_mode(4)
            mode = c.dense_mode()
            if mode in (3, 6):
                _capi.check(c.handle, L.clp_matvec_dev(c.handle, v.data_ptr(), 1.0, y.data_ptr(), None, None, 5, C.byref(ms)))
                _capi.check(c.handle, L.clp_matvec_dev(c.handle, v.data_ptr(), 1.0, y.data_ptr(), None, None, 50, C.byref(ms)))
                kept, pb = c.sparse_info()
                rec.update(compact_mode=mode, compact_ms=ms.value, compact_bytes=pb, compact_GBps=pb / ms.value / 1e6,
                           compact_frac=pb / ms.value / 1e6 / peak, compact_l2_resident=bool(pb < 126e6))
            emit(rec)
            del c
            torch.cuda.empty_cache()


def full_problem(name):
    prob = datagen.config_problem(name); cfg = prob["cfg"]; m = cfg["m"]
    if cfg["kind"] == "euclidean":
        ip = clp.invariants.EuclideanDistanceParams(); ip.sigma, ip.epsilon = cfg["sigma"], cfg["epsilon"]
        inv = clp.invariants.EuclideanDistance(ip)
    else:
        ip = clp.invariants.PointNormalDistanceParams()
        ip.sigp, ip.epsp, ip.sign, ip.epsn = cfg["sigp"], cfg["epsp"], cfg["sign"], cfg["epsn"]
        inv = clp.invariants.PointNormalDistance(ip)
    c = clp.CLIPPER(inv, clp.Params())
    ts, tv = [], []
    for rep in range(7):
        torch.cuda.synchronize()
        t0 = time.perf_counter(); c.score_pairwise_consistency(prob["D1"], prob["D2"], prob["A"]); t1 = time.perf_counter()
        c.solve(prob["u0"]); t2 = time.perf_counter()
        ts.append(t1 - t0); tv.append(t2 - t1)
    s = c.get_solution()
    nM, _ = c.count_nonzeros()
    no = m - prob["ni"]
    sel = np.asarray(s.nodes)
    tp = int((sel >= no).sum())
    info = oracle_step(prob, cpu_cores())
    emit({"config": name, "kind": cfg["kind"], "m": m, "density": nM / (m * (m - 1) / 2),
          "gpu_t_score_ms": 1e3 * float(np.median(ts[2:])), "gpu_t_solve_ms": 1e3 * float(np.median(tv[2:])),
          "solver_kernel_ms": s.kernel_ms, "evals": s.n_evals, "ifinal": s.ifinal, "F": s.score, "n_nodes": len(s.nodes),
          "precision": tp / max(1, len(sel)), "recall": tp / prob["ni"],
          "associations_per_s": m / float(np.median(ts[2:]) + np.median(tv[2:])),
          "cpu_oracle": {"cores": cpu_cores(), "t_score_s": info["t_score"], "t_solve_s": info["t_solve"], "evals": info["evals"],
                         "same_inlier_set": sorted(info["nodes"]) == sorted(s.nodes),
                         "rel_dF": abs(info["score"] - s.score) / abs(info["score"])}})


if __name__ == "__main__":
    which = sys.argv[1:] or ["c1", "c3", "c5"]
    if "c1" in which:
        full_problem("c1")
    if "c3" in which:
        full_problem("c3")
    if "c5" in which:
        matvec_sweep()
