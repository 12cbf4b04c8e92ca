run() { n=$1; tag=$2; shift 2; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 --no-config4 > gpurun_out/bench_$tag.log 2>&1; echo "$tag rc=$?"; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"solver_kernel_ms": [0-9.]*\|"solver_phase_ms": {[^}]*}' gpurun_out/bench_$tag.log | tr '\n' ' '; echo; }
run 8 n8 X=1
run 8 n8c3 CLP_CTAS_PER_SM=3
run 4 n4 X=1
