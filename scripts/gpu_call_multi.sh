#!/bin/bash
# one 8-GPU box: C4 / C5 at 8 GPUs, strong scaling of C3 (global batch 32 over 2 / 4 / 8 ranks), the 1-rank-vs-2-rank id test
mkdir -p gpurun_out
run() { name=$1; np=$2; port=$3; shift 3
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $port bench.py --gpus $np "$@" \
    > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; tail -c 600 gpurun_out/$name.json | head -c 300; echo; }
timeout 900 python -m pytest tests/test_gpu_dp_ids.py -q -m gpu > gpurun_out/t_dp.log 2>&1; echo "dp ids rc=$?"; tail -3 gpurun_out/t_dp.log
run r02_bench_c4_8gpu 8 29611 --workload c4 --steps 3 --warmup 3 --no-cpu-baseline
run r02_bench_c5_8gpu 8 29612 --workload c5 --steps 3 --warmup 3 --no-cpu-baseline
run r02_bench_c3_strong_g32_8gpu 8 29613 --workload c3 --global-batch 32 --steps 3 --warmup 3 --no-cpu-baseline
run r02_bench_c3_strong_g32_4gpu 4 29614 --workload c3 --global-batch 32 --steps 3 --warmup 3 --no-cpu-baseline
run r02_bench_c3_strong_g32_2gpu 2 29615 --workload c3 --global-batch 32 --steps 3 --warmup 3 --no-cpu-baseline
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02_bench_*gpu.json')):
    try:
        d = json.load(open(f)); print(f, d['n_gpus'], d['scaling'], round(d['value'], 2), d['unit'], round(d['ms_per_step'], 1), 'ms/step e2e', round(d['e2e']['value'], 2))
    except Exception as e: print(f, 'ERR', e)
PY
