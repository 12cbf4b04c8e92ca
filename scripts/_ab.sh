timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for f in 1 0; do CLP_SCORE_FILTER=$f timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/sf_$f.json 2>gpurun_out/sf_$f.err; done
python - <<'PY'
import json,glob
for n in ["gpurun_out/sf_1.json","gpurun_out/sf_0.json"]:
    try:
        j=json.loads(open(n).read().strip().splitlines()[-1])
        print(n,"value %.0f"%j["value"],"ms %.2f"%j["ms_per_step"],"e2e %.0f"%j["e2e"]["value"],"kernel %.2f"%j["config"].get("solver_kernel_ms"),{k:round(v,2) for k,v in j["config"].get("solver_phase_ms").items()},"mv %.3f"%j["config"]["matvec_alone_frac"],"roof %.3f"%j["roofline"]["frac"])
    except Exception as e: print(n,"ERR",e)
PY
