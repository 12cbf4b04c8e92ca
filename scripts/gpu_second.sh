#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
# launch list (every launch with its device time; serialised, cold cache: compare shares)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r01.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu list rc=$?"
# full capture of the dominant kernel (solver) and of the scoring + stand-alone mat-vec kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:solver_kernel -s 3 -c 1 -o gpurun_out/solver_r01 -f \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_solver.log 2>&1
echo "ncu solver rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"score_tile|matvec_partials" -s 4 -c 2 -o gpurun_out/score_matvec_r01 -f \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_score.log 2>&1
echo "ncu score rc=$?"
ls -la gpurun_out
