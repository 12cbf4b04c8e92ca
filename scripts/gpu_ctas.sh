#!/bin/bash
N=${1:-8}
for c in 1 2 3; do
  CLP_CTAS_PER_SM=$c timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $N --steps 5 --warmup 3 --no-config4 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); c=d['config']
print('N=$N ctas=$c value %.0f ms/step %.2f kernel %.2f %s'%(d['value'], d['ms_per_step'], c['solver_kernel_ms'], c['solver_phase_ms']))"
done
