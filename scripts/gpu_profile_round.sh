#!/bin/bash
# ncu evidence for the default configuration (auto sweep) of bench.py: launch list + full capture of the
# dominant kernel (persistent solver) + the scoring kernel + the sparse build kernels
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r01c.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:solver_kernel -s 3 -c 1 -o gpurun_out/solver_r01c -f \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_solver.log 2>&1
echo "solver rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"matvec_sparse_partials|sparse_fill|sparse_count" -s 6 -c 3 -o gpurun_out/sparse_r01c -f \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_sparse.log 2>&1
echo "sparse rc=$?"
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 3000 gpurun_out/bench_default.json
