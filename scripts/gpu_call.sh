#!/bin/bash
timeout -s KILL 19 python -m pytest tests/test_gpu_towers.py -q -x 2>&1 | grep -E "passed|failed|rror" | tail -2
