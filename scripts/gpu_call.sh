#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_eval_drivers.py tests/test_gpu_chanattn.py tests/test_gpu_gemm_pair.py -q 2>&1 | grep -E "passed|failed" | tail -2
( time timeout -s KILL 900 python bench.py > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err ) 2> gpurun_out/bench_time.log; grep real gpurun_out/bench_time.log
( time timeout -s KILL 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err ) 2> gpurun_out/ref_time.log; grep real gpurun_out/ref_time.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_c3.json'))
print('c3', round(d['value'],2), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],2), d['clocks'])
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['cpu_baseline']['stage_seconds'])
r=json.load(open('gpurun_out/r02_bench_reference.json'))
print('ref', r['value'], r['steps'], r['cpu_baseline']['cores'], r['cpu_baseline']['stage_seconds'])
PY
