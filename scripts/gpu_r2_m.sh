#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export CLP_SKIP_C4=1
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_batch.py -m gpu -q --maxfail=10 > gpurun_out/pytest_m.log 2>&1; tail -3 gpurun_out/pytest_m.log
for fl in 1 0; do
  CLP_STAGE_FLAGS=$fl CLP_PROF_CTAS=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-config4 > gpurun_out/benchm_f$fl.json 2> gpurun_out/benchm_f$fl.err
  python - $fl <<'PY'
import json,sys
f=sys.argv[1]
d=json.load(open("gpurun_out/benchm_f%s.json"%f)); c=d["config"]
print("flags=%s"%f, "value %.5g"%d["value"], "ms/step %.3f"%d["ms_per_step"], "solver %.3f"%c["solver_kernel_ms"], {k:round(v,3) for k,v in c["solver_phase_ms"].items()}, "roofline %.3f"%d["roofline"]["frac"], "e2e %.5g"%d["e2e"]["value"])
PY
  grep "clp prof" gpurun_out/benchm_f$fl.err | tail -4
done
timeout 300 python scripts/sweep.py c1 c3 2>&1 | tail -2 | cut -c1-330
