#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export CLP_SKIP_C4=1
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_gpu_batch.py tests/test_shell_and_pybind.py -m gpu -q --maxfail=10 > gpurun_out/pytest_e.log 2>&1
echo "pytest(e) rc=$?" >> gpurun_out/pytest_e.log
tail -4 gpurun_out/pytest_e.log
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
for cfg in 1 0; do
    CLP_PROF_CTAS=1 CLP_RES_CFG=$cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-config4 > gpurun_out/benche_cfg${cfg}.json 2> gpurun_out/benche_cfg${cfg}.err
    show "cfg$cfg" gpurun_out/benche_cfg${cfg}.json
    grep "clp prof" gpurun_out/benche_cfg${cfg}.err | tail -3
done
for G in 0 1 2 4 8 32 125; do
  CLP_RES_G=$G timeout 300 python scripts/sweep.py c1 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('c1 G=$G', 'solve %.3f ms'%d['gpu_t_solve_ms'], 'kernel %.3f ms'%d['solver_kernel_ms'], 'evals', d['evals'], 'us/eval %.2f'%(1e3*d['solver_kernel_ms']/d['evals']), 'assoc/s %.4g'%d['associations_per_s'])"
done
for m in 3000 6000; do
for G in 0 148; do
  CLP_RES_G=$G python - <<PY
import numpy as np, clipper_b200 as clp
from clipper_b200 import datagen
prob=datagen.config_problem("c2",$m); cfg=prob["cfg"]
ip=clp.invariants.EuclideanDistanceParams(); ip.sigma,ip.epsilon=cfg["sigma"],cfg["epsilon"]
c=clp.CLIPPER(clp.invariants.EuclideanDistance(ip),clp.Params())
c.score_pairwise_consistency(prob["D1"],prob["D2"],prob["A"])
for _ in range(3): c.solve(prob["u0"])
s=c.get_solution(); print("m=$m G=$G kernel %.3f ms evals %d us/eval %.2f prof %s"%(s.kernel_ms,s.n_evals,1e3*s.kernel_ms/s.n_evals,[round(x,3) for x in s.prof_ms]))
PY
done
done
