#!/bin/bash
mkdir -p gpurun_out
FO1_MEGA_PROF=1 timeout 600 python scripts/mega_prof.py 32 1195 > gpurun_out/mega_prof.log 2>&1; echo "rc=$?"; grep -E "decode_mega profile|decode ms" gpurun_out/mega_prof.log | tail -12
FO1_MEGA_PROF=1 timeout 600 python scripts/mega_prof.py 8 1195 > gpurun_out/mega_prof8.log 2>&1; grep -E "decode_mega profile|decode ms" gpurun_out/mega_prof8.log | tail -6
