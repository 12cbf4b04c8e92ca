#!/bin/bash
# full -m gpu suite including the m = 80000 oracle comparison; sanitizer runs on c1; c3 with the bulk-copy ring
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/fullsize_parity.jsonl
timeout 2400 python -m pytest tests -m gpu -q --maxfail=10 --durations=8 > gpurun_out/pytest_j.log 2>&1
echo "pytest(j) rc=$?" >> gpurun_out/pytest_j.log
tail -16 gpurun_out/pytest_j.log
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_c1.py > gpurun_out/sanitizer_${tool}_c1.log 2>&1
  echo "sanitizer $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|c1 mode|batch F" gpurun_out/sanitizer_${tool}_c1.log | tail -4
done
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_c1.py 3 > gpurun_out/sanitizer_memcheck_c1_mode3.log 2>&1; grep -E "ERROR SUMMARY|c1 mode" gpurun_out/sanitizer_memcheck_c1_mode3.log | tail -2
for cfg in 1 0 4 5 6; do
  CLP_RES_CFG=$cfg python - <<'PY'
import os, numpy as np, clipper_b200 as clp, ctypes as C, torch
from clipper_b200 import datagen, _capi
prob=datagen.config_problem("c3"); cfg=prob["cfg"]; m=cfg["m"]
ip=clp.invariants.PointNormalDistanceParams()
c=clp.CLIPPER(clp.invariants.PointNormalDistance(ip),clp.Params())
c.score_pairwise_consistency(prob["D1"],prob["D2"],prob["A"])
for _ in range(3): c.solve(prob["u0"])
s=c.get_solution()
L=_capi.load(); dev=torch.device("cuda:0")
v=torch.rand(m,dtype=torch.float64,device=dev); y=torch.empty_like(v); ms=C.c_double()
_capi.check(c.handle, L.clp_matvec_dev(c.handle, v.data_ptr(), 1.0, y.data_ptr(), None, None, 5, C.byref(ms)))
_capi.check(c.handle, L.clp_matvec_dev(c.handle, v.data_ptr(), 1.0, y.data_ptr(), None, None, 50, C.byref(ms)))
kept,pb=c.sparse_info()
print("c3 cfg",os.environ["CLP_RES_CFG"],"mode",c.dense_mode(),"solver %.3f ms evals %d F %.6f"%(s.kernel_ms,s.n_evals,s.score),"matvec alone %.4f ms %.0f GB/s frac %.3f"%(ms.value,pb/ms.value/1e6,pb/ms.value/1e6/6574.1), "prof",[round(x,3) for x in s.prof_ms])
PY
done
