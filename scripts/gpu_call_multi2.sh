#!/bin/bash
# one 2-GPU box: the 1-rank-vs-2-rank id test and the 2-rank point of the strong-scaling curve (C3, global batch 32)
mkdir -p gpurun_out
run() { name=$1; np=$2; port=$3; shift 3
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $port bench.py --gpus $np "$@" \
    > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; tail -c 600 gpurun_out/$name.json | head -c 300; echo; }
timeout 900 python -m pytest tests/test_gpu_dp_ids.py -q -m gpu > gpurun_out/t_dp.log 2>&1; echo "dp ids rc=$?"; tail -3 gpurun_out/t_dp.log
run r02_bench_c3_strong_g32_2gpu 2 29615 --workload c3 --global-batch 32 --steps 3 --warmup 3 --no-cpu-baseline
