#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 150 python -m pytest tests/test_gpu_attention.py -x -q 2>&1 | grep -E "passed|failed|Error|error" | tail -4
echo "--- persistent"; timeout -s KILL 120 python scripts/attn_probe.py 2>&1 | cut -c1-160 | tail -12
echo "--- one CTA per item"; FO1_ATTN_ONE_ITEM=1 timeout -s KILL 120 python scripts/attn_probe.py 2>&1 | cut -c1-160 | tail -12
