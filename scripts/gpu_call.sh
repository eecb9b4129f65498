#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_decode_mega.py tests/test_gpu_llm.py -x -q 2>&1 | tail -4
for cfg in "32 1195" "8 600"; do
set -- $cfg
FO1_MEGA_PROF=1 timeout -s KILL 600 python scripts/mega_prof.py $1 $2 > gpurun_out/mega_prof_$1_$2.log 2>&1; grep -E "ms/step|decode_mega profile, first" gpurun_out/mega_prof_$1_$2.log | tail -3
done
