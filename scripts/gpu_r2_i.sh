#!/bin/bash
# 2-GPU A/B of the LL staging load flavour; N=1 bench after the barrier fix
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export CLP_SKIP_C4=1
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_fullsize.py -m gpu -q --maxfail=10 > gpurun_out/pytest_i.log 2>&1; tail -3 gpurun_out/pytest_i.log
CLP_PROF_CTAS=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-config4 > gpurun_out/benchi.json 2> gpurun_out/benchi.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/benchi.json")); c=d["config"]
print("N=1", "value %.4g"%d["value"], "ms/step %.3f"%d["ms_per_step"], "solver %.3f"%c["solver_kernel_ms"], {k:round(v,3) for k,v in c["solver_phase_ms"].items()}, "roofline %.3f"%d["roofline"]["frac"])
PY
grep "clp prof" gpurun_out/benchi.err | tail -4
for scope in 1; do
  CLP_LL_GPU_SCOPE=$scope CLP_PROF_CTAS=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29700+scope)) bench.py --gpus 2 --steps 10 --warmup 3 --no-config4 > gpurun_out/benchi_n2_s$scope.json 2> gpurun_out/benchi_n2_s$scope.err
  echo "N=2 gpu_scope=$scope rc=$?"; grep "clp prof" gpurun_out/benchi_n2_s$scope.err | grep "148 CTAs" | tail -4
  python - $scope <<'PY'
import json,sys
s=sys.argv[1]
try:
    d=[json.loads(l) for l in open("gpurun_out/benchi_n2_s%s.json"%s) if l.startswith("{")][-1]; c=d["config"]
    print("N=2 scope=%s"%s, "value %.4g"%d["value"], "ms/step %.3f"%d["ms_per_step"], "solver %.3f"%c["solver_kernel_ms"], {k:round(v,3) for k,v in c["solver_phase_ms"].items()})
except Exception as e:
    print("failed", e); print(open("gpurun_out/benchi_n2_s%s.err"%s).read()[-2000:])
PY
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29710 scripts/check_sharded.py 3000 20000 2>&1 | grep '^{' | tail -2
