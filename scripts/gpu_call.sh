#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_hfre.py tests/test_gpu_dwconv.py tests/test_gpu_pipeline.py tests/test_gpu_boundary.py tests/test_gpu_towers.py -x -q 2>&1 | tail -5
timeout 600 python scripts/davit_prof.py 32 768 2>&1 | tail -1
timeout 900 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -2 gpurun_out/bench_c2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c2.json'))
print(d['value'], d['unit'], d['ms_per_step'], d['stage_ms'], d['clocks'])
print(d['roofline'])
PY
