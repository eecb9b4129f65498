#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:pair_kernel -s 1 -c 1 -o gpurun_out/r02_pair_gemm -f python scripts/ncu_pair.py > gpurun_out/ncu_pair.log 2>&1; tail -2 gpurun_out/ncu_pair.log
timeout -s KILL 1500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
  -k regex:'gemm_bf16|attn_tc|hfre_sweep|chan_gram|chan_apply|dwconv3x3' --csv --log-file gpurun_out/r02_dram_step.csv \
  python bench.py --profile-run --steps 1 --warmup 0 --tokens 1 --no-cpu-baseline > gpurun_out/r02_dram_step.log 2>&1
tail -1 gpurun_out/r02_dram_step.log; python scripts/make_traffic.py gpurun_out/r02_dram_step.csv gpurun_out/r02_traffic.json
( time timeout -s KILL 1500 python -m pytest tests -q -m gpu > gpurun_out/r02_pytest_gpu.log 2>&1 ) 2> gpurun_out/pytest_time.log; tail -2 gpurun_out/r02_pytest_gpu.log; grep real gpurun_out/pytest_time.log
