#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout 900 ncu --set full --clock-control none --import-source on -k regex:matvec2_partials -s 3 -c 1 -o gpurun_out/matvec2_sym_r01 -f \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_mv2.log 2>&1
echo "ncu rc=$?"
