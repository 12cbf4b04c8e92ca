#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export CLP_SKIP_C4=1
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_gpu_batch.py tests/test_shell_and_pybind.py -m gpu -q --maxfail=10 > gpurun_out/pytest_g.log 2>&1
echo "pytest(g) rc=$?" >> gpurun_out/pytest_g.log
tail -4 gpurun_out/pytest_g.log
CLP_PROF_HOST=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-config4 > gpurun_out/benchg.json 2> gpurun_out/benchg.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/benchg.json")); c=d["config"]
print("N=1", "value %.4g"%d["value"], "ms/step %.3f"%d["ms_per_step"], "solver %.3f"%c["solver_kernel_ms"], {k:round(v,3) for k,v in c["solver_phase_ms"].items()}, "roofline %.3f"%d["roofline"]["frac"], "e2e %.4g"%d["e2e"]["value"])
PY
grep "clp host" gpurun_out/benchg.err | sed -n 21,26p
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02g_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-config4 > gpurun_out/ncu_launches_g.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02g_launches.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
seq=[(r[ik].split('(')[0][:50], float(r[iv].replace(',',''))) for r in rows[1:]]
for k,v in seq[-10:]: print("%12.1f us  %s"%(v/1e3,k))
PY
