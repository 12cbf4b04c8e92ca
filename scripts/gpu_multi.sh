#!/bin/bash
# usage: gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x > gpurun_out/pytest_sharded_$N.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_sharded_$N.log
tail -8 gpurun_out/pytest_sharded_$N.log
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
       bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_n$n.log 2>&1
    echo "bench n=$n rc=$?"; tail -c 2500 gpurun_out/bench_n$n.log
  fi
done
