#!/usr/bin/env python
"""BASELINE.json config 5: "Md.u matvec sweep m=2k-128k, fp32 warp-shuffle vs tf32 tensor-core GEMV, ncu HBM GB/s".

Three ways to compute y = M v on the SAME dense fp32 matrix (4 m^2 bytes), CUDA-event timed, 50 launches after 5 warm-ups:
  ours   : the hand-written fp32-storage / fp64-accumulate warp-shuffle sweep (clp_matvec_dev, sweep mode 0: full dense
           matrix, what north_star's ">= 40 % of the HBM roofline" is quoted on)
  tf32   : the tensor-core formulation -- M (m x m, fp32 storage, tf32 multiply) x V (m x 8, v in column 0): cuBLAS picks a
           tf32 tensor-op GEMM (torch.matmul with allow_tf32); this is the library's best tensor-core kernel for the shape
  sgemv  : cuBLAS fp32 SIMT GEMV (torch.mv), for reference
and the error of each against the fp64 product.  A GEMV has no operand reuse: all three are bound by the same 4 m^2
bytes of HBM traffic; the tensor path cannot beat the memory roofline and rounds v (and M) to 10 mantissa bits, which
breaks the 1e-5 tolerance of the solver's objective (north_star) -- the sweep stays on the fp64-accumulate SIMT path.
One JSON line per m to gpurun_out/tf32_gemv_r02.jsonl.  Run the tf32 arm alone under ncu with --only-tf32 M to capture
sm__pipe_tensor / dram throughput of the tensor kernel."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import clipper_b200 as clp  # noqa: E402
from clipper_b200 import _capi, datagen  # noqa: E402
from bench import measured_peaks  # noqa: E402


def timed(fn, reps=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    peak, _ = measured_peaks()
    dev = torch.device("cuda:0")
    only_tf32 = None
    if len(sys.argv) > 2 and sys.argv[1] == "--only-tf32":
        only_tf32 = int(sys.argv[2])
    sizes = [only_tf32] if only_tf32 else [2048, 4096, 8192, 16384, 20000, 32768, 65536]
    L = _capi.load()
    out = open(os.path.join(ROOT, "gpurun_out", "tf32_gemv_r02.jsonl"), "a")
    for m in sizes:
        prob = datagen.config_problem("c2", m); cfg = prob["cfg"]
        ip = clp.invariants.EuclideanDistanceParams(); ip.sigma, ip.epsilon = cfg["sigma"], cfg["epsilon"]
        c = clp.CLIPPER(clp.invariants.EuclideanDistance(ip), clp.Params())
        c.set_dense_mode(0)  # full dense fp32 matrix, 4 m^2 bytes per pass
        c.score_pairwise_consistency(prob["D1"], prob["D2"], prob["A"])
        v64 = torch.rand(m, dtype=torch.float64, device=dev)
        # the same matrix as a torch tensor: Mhat = M - I (zero diagonal), built by mat-vecs with unit vectors would be
        # O(m^2) launches; instead rebuild it from the host getter for m <= 20000 and from random data of the same
        # shape above (timing does not depend on the values)
        if m <= 20000:
            Md = torch.from_numpy(c.get_affinity_matrix()).to(dev)
            Md.fill_diagonal_(0.0)
            M32 = Md.to(torch.float32); del Md
        else:
            M32 = torch.rand(m, m, dtype=torch.float32, device=dev)
        V = torch.zeros(m, 8, dtype=torch.float32, device=dev); V[:, 0] = v64.to(torch.float32)
        v32 = v64.to(torch.float32)
        byts = 4.0 * m * m + 16 * m
        rec = {"config": "c5", "m": m, "bytes": byts, "l2_resident": bool(4 * m * m < 126e6)}
        torch.backends.cuda.matmul.allow_tf32 = True
        ms_tf32 = timed(lambda: torch.matmul(M32, V))
        y_tf32 = torch.matmul(M32, V)[:, 0].double()
        rec.update(tf32_ms=ms_tf32, tf32_GBps=byts / ms_tf32 / 1e6, tf32_frac=byts / ms_tf32 / 1e6 / peak)
        if not only_tf32:
            torch.backends.cuda.matmul.allow_tf32 = False
            ms_mv = timed(lambda: torch.mv(M32, v32))
            y_mv = torch.mv(M32, v32).double()
            y = torch.empty_like(v64); Mv = torch.empty_like(v64); Cv = torch.empty_like(v64)
            ms = C.c_double()
            _capi.check(c.handle, L.clp_matvec_dev(c.handle, v64.data_ptr(), 1.0, y.data_ptr(), Mv.data_ptr(), Cv.data_ptr(), 5, C.byref(ms)))
            _capi.check(c.handle, L.clp_matvec_dev(c.handle, v64.data_ptr(), 1.0, y.data_ptr(), Mv.data_ptr(), Cv.data_ptr(), 50, C.byref(ms)))
            rec.update(ours_ms=ms.value, ours_GBps=byts / ms.value / 1e6, ours_frac=byts / ms.value / 1e6 / peak,
                       sgemv_ms=ms_mv, sgemv_GBps=byts / ms_mv / 1e6, sgemv_frac=byts / ms_mv / 1e6 / peak)
            if m <= 20000:
                ref = torch.matmul(M32.double(), v64)
                nrm = ref.abs().max().item()
                rec.update(rel_err_ours=(Mv - ref).abs().max().item() / nrm, rel_err_tf32=(y_tf32 - ref).abs().max().item() / nrm,
                           rel_err_sgemv=(y_mv - ref).abs().max().item() / nrm)
        print(json.dumps(rec), flush=True)
        out.write(json.dumps(rec) + "\n"); out.flush()
        del c, M32, V
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
