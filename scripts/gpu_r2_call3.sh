#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 --timeout-method thread 2>&1 | tail -4
bash scripts/gpu_stats.sh > gpurun_out/tc_stats.txt 2>&1
cat gpurun_out/tc_stats.txt
