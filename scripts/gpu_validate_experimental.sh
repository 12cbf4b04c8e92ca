#!/bin/bash
# One gpurun call that validates and measures everything that was written after round 1's GPU budget ran out:
#   - clp_set_grid_cap / BatchSolver (tests/test_gpu_batch.py, scripts/bench_batch.py)
#   - the cp.async ring sweep solver_kernel<float,5> (CLP_SPARSE_RING=<depth>)
# usage (single GPU):  gpurun --timeout 900 -- 'bash scripts/gpu_validate_experimental.sh'
mkdir -p gpurun_out
CLP_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_batch.py -q -m gpu 2>&1 | tail -15 | tee gpurun_out/exp_tests.log
# multi-GPU (run under gpurun --gpus N): CLP_SHARD_INTERLEAVE=1 CLP_BENCH_E2E_MULTI=1 python -m torch.distributed.run ... bench.py --gpus N
for d in 0 2 3 4 6; do
  CLP_SPARSE_RING=$d timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/exp_ring_$d.json 2> gpurun_out/exp_ring_$d.err
done
timeout 300 python scripts/bench_batch.py 1000 256 18 24 > gpurun_out/exp_batch_m1000.json 2> gpurun_out/exp_batch_m1000.err
timeout 300 python scripts/bench_batch.py 1000 256 36 12 > gpurun_out/exp_batch_m1000_cap12.json 2> gpurun_out/exp_batch_m1000_cap12.err
timeout 300 python scripts/bench_batch.py 2000 128 12 36 > gpurun_out/exp_batch_m2000.json 2> gpurun_out/exp_batch_m2000.err
python - <<'PY'
import json, glob
for n in sorted(glob.glob("gpurun_out/exp_ring_*.json")):
    try:
        j = json.loads(open(n).read().strip().splitlines()[-1])
        print(n, "value %.0f" % j["value"], "ms %.2f" % j["ms_per_step"], "kernel %.2f" % j["config"]["solver_kernel_ms"],
              {k: round(v, 2) for k, v in j["config"]["solver_phase_ms"].items()})
    except Exception as e:
        print(n, "ERR", e, open(n.replace(".json", ".err")).read()[-400:])
for n in sorted(glob.glob("gpurun_out/exp_batch_*.json")):
    try:
        print(n, open(n).read().strip().splitlines()[-1])
    except Exception as e:
        print(n, "ERR", e, open(n.replace(".json", ".err")).read()[-400:])
PY
