#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CLP_PROF_CTAS=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-config4 > gpurun_out/benchf.json 2> gpurun_out/benchf.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/benchf.json")); c=d["config"]
print("N=1", "value %.4g"%d["value"], "ms/step %.3f"%d["ms_per_step"], "solver %.3f"%c["solver_kernel_ms"], {k:round(v,3) for k,v in c["solver_phase_ms"].items()}, "roofline %.3f"%d["roofline"]["frac"])
PY
grep "clp prof" gpurun_out/benchf.err | tail -3
bash scripts/gpu_r2_multi.sh
