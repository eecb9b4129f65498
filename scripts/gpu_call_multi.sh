#!/bin/bash
# one 8-GPU box: C4 / C5 at 8 GPUs (weak), strong scaling of C3 (global batch 32 over 8 and 4 ranks)
mkdir -p gpurun_out
run() { name=$1; np=$2; port=$3; shift 3
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $port bench.py --gpus $np "$@" \
    > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; tail -c 600 gpurun_out/$name.json | head -c 300; echo; }
run r02_bench_c4_8gpu 8 29611 --workload c4 --steps 3 --warmup 3 --no-cpu-baseline
run r02_bench_c5_8gpu 8 29612 --workload c5 --steps 3 --warmup 3 --no-cpu-baseline
run r02_bench_c3_strong_g32_8gpu 8 29613 --workload c3 --global-batch 32 --steps 3 --warmup 3 --no-cpu-baseline
run r02_bench_c3_strong_g32_4gpu 4 29614 --workload c3 --global-batch 32 --steps 3 --warmup 3 --no-cpu-baseline
