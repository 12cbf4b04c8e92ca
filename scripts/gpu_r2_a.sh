#!/bin/bash
# round-2 first GPU pass: parity tests (c4 skipped here), then A/B of the resident solver's load pipelines at c2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export CLP_SKIP_C4=1
timeout 1200 python -m pytest tests -m gpu -q --maxfail=12 -x --deselect tests/test_gpu_fullsize.py > gpurun_out/pytest_a.log 2>&1
echo "pytest(a) rc=$?" >> gpurun_out/pytest_a.log
tail -5 gpurun_out/pytest_a.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q --maxfail=12 > gpurun_out/pytest_full.log 2>&1
echo "pytest(full) rc=$?" >> gpurun_out/pytest_full.log
tail -5 gpurun_out/pytest_full.log
for cfg in 0 1 2 3 4 5 6; do
  CLP_RES_CFG=$cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg$cfg.json 2> gpurun_out/bench_cfg$cfg.err
  echo "cfg $cfg rc=$?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_cfg$cfg.json")); c=d["config"]
    print("cfg $cfg", "value %.4g"%d["value"], "ms/step %.3f"%d["ms_per_step"], "solver %.3f"%c["solver_kernel_ms"], c["solver_phase_ms"], "mv alone frac %.3f"%c["matvec_alone_frac"], "roofline %.3f"%d["roofline"]["frac"], "e2e %.4g"%d["e2e"]["value"])
except Exception as e:
    print("cfg $cfg failed", e); print(open("gpurun_out/bench_cfg$cfg.err").read()[-2000:])
PY
done
CLP_RESIDENT=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_old.json 2> gpurun_out/bench_old.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_old.json")); c=d["config"]
    print("old path", "value %.4g"%d["value"], "ms/step %.3f"%d["ms_per_step"], "solver %.3f"%c["solver_kernel_ms"], c["solver_phase_ms"], "mv alone frac %.3f"%c["matvec_alone_frac"])
except Exception as e:
    print("old failed", e); print(open("gpurun_out/bench_old.err").read()[-2000:])
PY
