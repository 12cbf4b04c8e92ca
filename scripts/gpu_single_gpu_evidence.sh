#!/bin/bash
# round-2 final single-GPU evidence: driver-style bench (+ reference arm sanity), smoke, ncu, sweeps
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02l_bench_default.json 2> gpurun_out/r02l_bench_default.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open("gpurun_out/r02l_bench_default.json")); c=d["config"]
print("value %.5g"%d["value"], "ms/step %.3f"%d["ms_per_step"], "e2e %.5g"%d["e2e"]["value"], "solver %.3f"%c["solver_kernel_ms"], {k:round(v,3) for k,v in c["solver_phase_ms"].items()}, "roofline", {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["roofline"].items() if k in ("achieved","frac","traffic")})
print("cpu_baseline", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d.get("cpu_baseline",{}).items() if k!="sample"})
print("config4", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d.get("config4",{}).items() if k!="note"})
print("clocks", d["clocks"], "dense_matvec_alone", c.get("dense_matvec_alone"))
PY
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/r02l_bench_reference.json 2>&1; cut -c1-700 gpurun_out/r02l_bench_reference.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:solver_resident -s 3 -c 1 -o gpurun_out/r02l_resident -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-config4 > gpurun_out/ncu_resident_l.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02l_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-config4 > gpurun_out/ncu_launches_l.log 2>&1; echo "ncu launches rc=$?"
timeout 1500 python scripts/sweep.py c1 c3 c5 > gpurun_out/sweep_l.log 2>&1; tail -3 gpurun_out/sweep_l.log | cut -c1-500
