#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_preprocess.py tests/test_gpu_prompt.py -q -m gpu > gpurun_out/t_new.log 2>&1; echo "new rc=$?"; grep -E "passed|failed|Error|assert" gpurun_out/t_new.log | tail -20
timeout 1500 python -m pytest tests/test_gpu_parity_stages.py -q -m gpu -s > gpurun_out/t_parity.log 2>&1; echo "parity rc=$?"; grep -E "passed|failed|Error|assert" gpurun_out/t_parity.log | tail -30
timeout 600 python -m pytest tests/test_gpu_boundary.py tests/test_gpu_towers.py -q -m gpu > gpurun_out/t_b.log 2>&1; echo "boundary rc=$?"; tail -5 gpurun_out/t_b.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench c3 rc=$?"; tail -3 gpurun_out/bench_c3.err
timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"; tail -3 gpurun_out/bench_c2.err
timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench c4 rc=$?"; tail -3 gpurun_out/bench_c4.err
timeout 600 python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "bench c5 rc=$?"; tail -3 gpurun_out/bench_c5.err
python - <<'PY'
import json
for w in ("c3","c2","c4","c5"):
    try:
        d=json.load(open(f'gpurun_out/bench_{w}.json'))
        print(w, {k:d.get(k) for k in ('value','unit','ms_per_step','gpu_launches','images_per_s','images_per_s_e2e')}, d.get('stage_ms'))
        print('  e2e', d.get('e2e'))
        for k in ('roofline','roofline_attn','roofline_attn_causal','roofline_hfre','roofline_decode'):
            r=d.get(k)
            if r: print('  ',k,{x:r[x] for x in ('achieved','frac','share_of_step','operator_ms','operator_frac','ms_per_step') if x in r})
    except Exception as e: print(w,'ERR',e)
PY
