#!/bin/bash
# A/B of the trial-vector staging of the resident solver: cp.async.bulk (CLP_STAGE_BULK=1, default) vs register loads (0)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export CLP_SKIP_C4=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_shell_and_pybind.py -m gpu -q --maxfail=5 2>&1 | tail -4
for B in 1 0 1 0; do
  CLP_STAGE_BULK=$B CLP_PROF_CTAS=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_bulk$B.json 2> gpurun_out/bench_bulk$B.err
  python - $B <<'PY'
import json,sys
B=sys.argv[1]
d=json.load(open("gpurun_out/bench_bulk%s.json"%B)); c=d["config"]
print("bulk=%s"%B, "value %.4g"%d["value"], "ms/step %.3f"%d["ms_per_step"], "solver %.3f"%c["solver_kernel_ms"], "frac %.3f"%d["roofline"]["frac"], "F", c["F"], c["n_nodes"], c["evals_per_solve"])
PY
  grep "clp prof" gpurun_out/bench_bulk$B.err | grep -i "staging" | tail -1
done
