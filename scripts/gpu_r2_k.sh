#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
show() {
python - "$1" "$2" <<'PY'
import json,sys
tag,f=sys.argv[1],sys.argv[2]
try:
    d=json.load(open(f)); c=d["config"]
    print(tag, "value %.4g"%d["value"], "ms/step %.3f"%d["ms_per_step"], "solver %.3f"%c["solver_kernel_ms"], {k:round(v,3) for k,v in c["solver_phase_ms"].items()},
          "mv alone %.3f ms frac %.3f"%(c["matvec_alone_ms"],c["matvec_alone_frac"]), "roofline %.3f"%d["roofline"]["frac"], "e2e %.4g"%d["e2e"]["value"])
except Exception as e:
    print(tag, "failed", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
}
for laps in 1 0; do
  CLP_PROF_LAPS=$laps timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-config4 > gpurun_out/benchk_laps$laps.json 2> gpurun_out/benchk_laps$laps.err
  show "laps=$laps" gpurun_out/benchk_laps$laps.json
done
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_c1.py > gpurun_out/sanitizer_${tool}_c1.log 2>&1
  echo "sanitizer $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|c1 sweep|batch F" gpurun_out/sanitizer_${tool}_c1.log | tail -4
done
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_c1.py 3 > gpurun_out/sanitizer_memcheck_c1_mode3.log 2>&1; grep -E "ERROR SUMMARY|c1 sweep" gpurun_out/sanitizer_memcheck_c1_mode3.log | tail -2
CLP_PROF_BATCH=1 timeout 900 python scripts/bench_batch.py 64,256,1024,2048 444 > gpurun_out/batch_bench_k.jsonl 2> gpurun_out/batch_bench_k.err
cat gpurun_out/batch_bench_k.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print({k:(round(v,1) if isinstance(v,float) else v) for k,v in d.items() if k in ('m','problems','batch_problems_per_s','batch_kernel_problems_per_s','batch_call_problems_per_s','cpu_oracle_problems_per_s','speedup_vs_cpu_box','same_as_oracle','single_path_loop_problems_per_s')})"
grep "batch prof" gpurun_out/batch_bench_k.err | awk 'NR%2==0' | tail -4
