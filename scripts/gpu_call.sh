#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python scripts/gemm_pair_bench.py 2>&1 | tail -12
timeout -s KILL 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -2 gpurun_out/bench_c3.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c3.json'))
print(d['value'], d['ms_per_step'], d['stage_ms'], d['e2e']['value'], d['clocks'])
print(d['roofline'])
PY
