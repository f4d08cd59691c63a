#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -x --timeout 600 2>&1 | tail -3
for wl in C "D1024 --rows 20000000" B1024 "E1024 --rows 100000000"; do timeout 300 python scripts/bench_summary.py $wl --steps 5 --warmup 3; done
