#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:pair_kernel -s 1 -c 1 -o gpurun_out/r02_pair_smallk -f python scripts/ncu_pair_smallk.py > gpurun_out/ncu_pair_smallk.log 2>&1; tail -1 gpurun_out/ncu_pair_smallk.log
