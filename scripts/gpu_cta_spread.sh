#!/bin/bash
# per-CTA phase times of the resident solver at c2 (one line per CTA) for different item costs of the byte-balanced partition
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cost in 16 64 160; do
  CLP_ITEM_COST=$cost CLP_PROF_CTAS=2 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-config4 > gpurun_out/spread_c$cost.json 2> gpurun_out/spread_c$cost.err
  python - $cost <<'PY'
import json,sys,re
import numpy as np
cost=sys.argv[1]
d=json.load(open("gpurun_out/spread_c%s.json"%cost)); c=d["config"]
print("item cost", cost, "value %.5g"%d["value"], "solver %.3f"%c["solver_kernel_ms"], {k:round(v,3) for k,v in c["solver_phase_ms"].items()})
rows=[list(map(float,l.split()[2:])) for l in open("gpurun_out/spread_c%s.err"%cost) if l.startswith("[clp cta]")]
a=np.array(rows[-148:])  # last solve
sw,ep,ex,st,items,chunks=a[:,1],a[:,2],a[:,3],a[:,4],a[:,5],a[:,6]
print("  sweeps min/mean/max %.3f %.3f %.3f | items min/mean/max %d %.1f %d | chunks spread %.4f"%(sw.min(),sw.mean(),sw.max(),items.min(),items.mean(),items.max(),(chunks.max()-chunks.min())/chunks.mean()))
A=np.stack([chunks,items,np.ones_like(items)],1); coef,res,_,_=np.linalg.lstsq(A,sw,rcond=None)
print("  fit sweep_ms = %.3e*chunks + %.3e*items + %.3f  -> one item costs %.1f chunks; corr(sweep,items)=%.3f corr(sweep,bid)=%.3f"%(coef[0],coef[1],coef[2],coef[1]/coef[0],np.corrcoef(sw,items)[0,1],np.corrcoef(sw,a[:,0])[0,1]))
print("  epilogue vs items corr %.3f; slowest 5 CTAs:"%np.corrcoef(ep,items)[0,1], [(int(a[i,0]),round(sw[i],3),int(items[i])) for i in np.argsort(-sw)[:5]], "fastest 5:", [(int(a[i,0]),round(sw[i],3),int(items[i])) for i in np.argsort(sw)[:5]])
PY
done
