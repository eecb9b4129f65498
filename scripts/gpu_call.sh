#!/bin/bash
for i in 1 2; do r=$(timeout -s KILL 60 python -m pytest tests/test_gpu_eval_drivers.py -x -q -k generate_batch_equals 2>&1 | grep -E "passed|failed" | tail -1 | cut -c1-10); echo "default(per-kernel) run $i: $r"; done
