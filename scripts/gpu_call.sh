#!/bin/bash
mkdir -p gpurun_out
( time timeout -s KILL 1200 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err ) 2> gpurun_out/ref_time.log; grep real gpurun_out/ref_time.log; tail -2 gpurun_out/r02_bench_reference.err | cut -c1-300
( time timeout -s KILL 1200 python bench.py > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err ) 2> gpurun_out/bench_time.log; grep real gpurun_out/bench_time.log
python - <<'PY'
import json
r=json.load(open('gpurun_out/r02_bench_reference.json'))
print('ref', r['value'], 'steps', r['steps'], 'warm', r['warmup'], 'ms/step', r['ms_per_step'], r['cpu_baseline']['cores'], r['cpu_baseline']['stage_seconds'])
print(r['cpu_baseline']['sample'][:400])
d=json.load(open('gpurun_out/r02_bench_c3.json'))
print('c3', round(d['value'],2), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],2), d['clocks'])
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['cpu_baseline']['stage_seconds'])
PY
