timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/v3.json 2>gpurun_out/v3.err
CLP_FILL_ITEMS=0 CLP_FUSE_COUNT=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/v3_old.json 2>gpurun_out/v3_old.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r01g.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
python - <<'PY'
import json,glob
for n in ["gpurun_out/v3.json","gpurun_out/v3_old.json"]:
    try:
        j=json.loads(open(n).read().strip().splitlines()[-1])
        print(n,"value %.0f"%j["value"],"ms %.2f"%j["ms_per_step"],"e2e %.0f"%j["e2e"]["value"],"kernel %.2f"%j["config"].get("solver_kernel_ms"),{k:round(v,2) for k,v in j["config"].get("solver_phase_ms").items()},"mv %.3f"%j["config"]["matvec_alone_frac"],"roof %.3f"%j["roofline"]["frac"])
    except Exception as e: print(n,"ERR",e)
PY
