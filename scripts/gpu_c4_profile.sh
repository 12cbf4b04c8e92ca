#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python scripts/c4_once.py 2>&1 | tail -2
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:solver_kernel -s 1 -c 1 -o gpurun_out/r02q_c4_solver -f python scripts/c4_once.py > gpurun_out/ncu_c4.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_c4.log
