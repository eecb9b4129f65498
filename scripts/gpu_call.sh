#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_towers.py -x -q 2>&1 | grep -E "assert|Error|passed|failed|^E " | head -20
