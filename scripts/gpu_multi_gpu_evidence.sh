#!/bin/bash
# multi-GPU pass: NGPU = number of GPUs of the box; sharded-vs-oracle check, two-device pytest, bench at N = 2, 4, 8
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
echo "GPUs: $NG"
export CLP_SKIP_C4=1
timeout 600 python -m pytest tests/test_gpu_sharded.py "tests/test_gpu_parity.py::test_error_behaviour" -m gpu -q > gpurun_out/pytest_multi.log 2>&1; tail -3 gpurun_out/pytest_multi.log
for N in ${NLIST:-2 4 8}; do
  if [ $N -le $NG ]; then
    if [ $N -eq $NG ] || [ "$CHECK_ALL" = "1" ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) scripts/check_sharded.py 3000 20000 > gpurun_out/check_sharded_n$N.log 2>&1
    echo "check N=$N rc=$?"; grep '^{' gpurun_out/check_sharded_n$N.log | tail -2; grep -i "error\|Traceback" gpurun_out/check_sharded_n$N.log | head -5
    fi
    CLP_PROF_CTAS=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
    echo "bench N=$N rc=$?"; grep "clp prof" gpurun_out/bench_n$N.err | grep "148 CTAs" | tail -4
    python - $N <<'PY'
import json,sys
N=sys.argv[1]
try:
    d=[json.loads(l) for l in open("gpurun_out/bench_n%s.json"%N) if l.startswith("{")][-1]; c=d["config"]
    print("N=%s"%N, "value %.4g"%d["value"], "ms/step %.3f"%d["ms_per_step"], "solver %.3f"%c["solver_kernel_ms"], {k:round(v,3) for k,v in c["solver_phase_ms"].items()}, "mode", c["sweep_mode"], "e2e", {k:(round(v,1) if isinstance(v,float) else v) for k,v in d["e2e"].items() if k!="note"})
    if "config4" in d: print("   config4:", {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["config4"].items()})
except Exception as e:
    print("N=%s failed"%N, e); print(open("gpurun_out/bench_n%s.err"%N).read()[-2500:])
PY
  fi
done
