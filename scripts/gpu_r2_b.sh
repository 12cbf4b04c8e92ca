#!/bin/bash
# round-2 GPU pass B: all parity tests (c4 skipped), pipeline A/B at c2, batch benchmark, ncu of the resident solver
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export CLP_SKIP_C4=1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/pytest_b.log 2>&1
echo "pytest(b) rc=$?" >> gpurun_out/pytest_b.log
tail -15 gpurun_out/pytest_b.log
show() {
python - "$1" "$2" <<'PY'
import json,sys
tag,f=sys.argv[1],sys.argv[2]
try:
    d=json.load(open(f)); c=d["config"]
    print(tag, "value %.4g"%d["value"], "ms/step %.3f"%d["ms_per_step"], "solver %.3f"%c["solver_kernel_ms"], {k:round(v,3) for k,v in c["solver_phase_ms"].items()},
          "evals %d"%c["evals_per_solve"], "mv alone %.3f ms frac %.3f"%(c["matvec_alone_ms"],c["matvec_alone_frac"]), "roofline %.3f"%d["roofline"]["frac"], "e2e %.4g"%d["e2e"]["value"])
except Exception as e:
    print(tag, "failed", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
}
for cfg in 0 1 2 3 4 5 6; do
  CLP_RES_CFG=$cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg$cfg.json 2> gpurun_out/bench_cfg$cfg.err
  show "cfg$cfg" gpurun_out/bench_cfg$cfg.json
done
CLP_RESIDENT=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_old.json 2> gpurun_out/bench_old.err
show old gpurun_out/bench_old.json
timeout 900 python scripts/bench_batch.py 64,256,1024,2048 256 > gpurun_out/batch_bench.jsonl 2> gpurun_out/batch_bench.err
cat gpurun_out/batch_bench.jsonl; tail -5 gpurun_out/batch_bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:solver_resident -s 3 -c 1 -o gpurun_out/r02b_resident -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_resident.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_resident.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
