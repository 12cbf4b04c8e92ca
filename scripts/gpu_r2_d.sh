#!/bin/bash
# round-2 GPU pass D: LL-cell exchange + on-chip epilogue tables: parity, A/B, c1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export CLP_SKIP_C4=1
timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/pytest_d.log 2>&1
echo "pytest(d) rc=$?" >> gpurun_out/pytest_d.log
tail -6 gpurun_out/pytest_d.log
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
for cfg in 0 1; do
  for extra in 1 0; do
    CLP_PROF_CTAS=1 CLP_RES_CFG=$cfg CLP_RES_SMEM_EXTRA=$extra timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-config4 > gpurun_out/benchd_cfg${cfg}_x$extra.json 2> gpurun_out/benchd_cfg${cfg}_x$extra.err
    show "cfg$cfg extra=$extra" gpurun_out/benchd_cfg${cfg}_x$extra.json
    grep "clp prof" gpurun_out/benchd_cfg${cfg}_x$extra.err | tail -3
  done
done
timeout 600 python scripts/sweep.py c1 > gpurun_out/sweep_c1_d.log 2>&1; tail -1 gpurun_out/sweep_c1_d.log
