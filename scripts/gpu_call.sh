#!/bin/bash
# one gpurun call: attention bring-up (tight timeouts: a protocol bug must not hang the box)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 420 python -m pytest tests/test_gpu_attention.py -q -x -m gpu -k "attention_varlen or rescale or strided" > gpurun_out/t_attn.log 2>&1
rc=$?; echo "attn rc=$rc"; tail -15 gpurun_out/t_attn.log
if [ $rc -eq 0 ]; then
  timeout 300 python scripts/attn_probe.py > gpurun_out/attn_probe.jsonl 2> gpurun_out/attn_probe.err; echo "probe rc=$?"; cat gpurun_out/attn_probe.jsonl
  timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1; echo "all rc=$?"; tail -15 gpurun_out/t_all.log
fi
