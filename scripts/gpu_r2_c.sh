#!/bin/bash
# round-2 GPU pass C: on-chip epilogue tables A/B, per-CTA spreads, batch phase profile, c1/c3/c5 sweeps, tf32 comparison
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export CLP_SKIP_C4=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_gpu_batch.py -m gpu -q --maxfail=10 > gpurun_out/pytest_c.log 2>&1
echo "pytest(c) rc=$?" >> gpurun_out/pytest_c.log
tail -6 gpurun_out/pytest_c.log
show() {
python - "$1" "$2" <<'PY'
import json,sys
tag,f=sys.argv[1],sys.argv[2]
try:
    d=json.load(open(f)); c=d["config"]
    print(tag, "value %.4g"%d["value"], "ms/step %.3f"%d["ms_per_step"], "solver %.3f"%c["solver_kernel_ms"], {k:round(v,3) for k,v in c["solver_phase_ms"].items()},
          "evals %d"%c["evals_per_solve"], "mv alone %.3f ms frac %.3f"%(c["matvec_alone_ms"],c["matvec_alone_frac"]), "roofline %.3f"%d["roofline"]["frac"], "e2e %.4g"%d["e2e"]["value"])
    if "config4" in d: print("   config4:", {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["config4"].items() if k!="note"})
except Exception as e:
    print(tag, "failed", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
}
for cfg in 0 1; do
  for extra in 1 0; do
    CLP_PROF_CTAS=1 CLP_RES_CFG=$cfg CLP_RES_SMEM_EXTRA=$extra timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-config4 > gpurun_out/benchc_cfg${cfg}_x$extra.json 2> gpurun_out/benchc_cfg${cfg}_x$extra.err
    show "cfg$cfg extra=$extra" gpurun_out/benchc_cfg${cfg}_x$extra.json
    grep "clp prof" gpurun_out/benchc_cfg${cfg}_x$extra.err | tail -3
  done
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/benchc_default.json 2> gpurun_out/benchc_default.err
show "default(+config4)" gpurun_out/benchc_default.json
CLP_PROF_BATCH=1 timeout 900 python scripts/bench_batch.py 256,1024,2048 444 > gpurun_out/batch_bench_c.jsonl 2> gpurun_out/batch_bench_c.err
cat gpurun_out/batch_bench_c.jsonl; grep "batch prof" gpurun_out/batch_bench_c.err | tail -8; grep -v "batch prof" gpurun_out/batch_bench_c.err | tail -5
timeout 900 python scripts/sweep.py c1 c3 > gpurun_out/sweep_c1c3.log 2>&1; tail -3 gpurun_out/sweep_c1c3.log
rm -f gpurun_out/tf32_gemv_r02.jsonl
timeout 900 python scripts/tf32_gemv_compare.py > gpurun_out/tf32.log 2>&1; tail -8 gpurun_out/tf32.log
timeout 600 ncu --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,gpu__time_duration.sum --clock-control none -k regex:"gemm|cutlass|sm100|sm90|sm80|xmma|tensorop" -c 3 --csv --log-file gpurun_out/r02c_tf32_gemm_ncu.csv python scripts/tf32_gemv_compare.py --only-tf32 20000 > gpurun_out/tf32_ncu.log 2>&1
echo "ncu tf32 rc=$?"; tail -4 gpurun_out/r02c_tf32_gemm_ncu.csv
