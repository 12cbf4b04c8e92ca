#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q > gpurun_out/pytest_o.log 2>&1; tail -2 gpurun_out/pytest_o.log
CLP_PROF_CTAS=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29720 bench.py --gpus 2 --steps 10 --warmup 3 --no-config4 > gpurun_out/bencho_n2.json 2> gpurun_out/bencho_n2.err
grep "clp prof" gpurun_out/bencho_n2.err | grep "148 CTAs" | tail -4
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/bencho_n2.json") if l.startswith("{")][-1]; c=d["config"]
print("N=2", "value %.4g"%d["value"], "ms/step %.3f"%d["ms_per_step"], "solver %.3f"%c["solver_kernel_ms"], {k:round(v,3) for k,v in c["solver_phase_ms"].items()})
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29721 scripts/check_sharded.py 20000 2>&1 | grep '^{' | tail -1
