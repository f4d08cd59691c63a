#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -q -x --timeout 600 2>&1 | tail -3
timeout 300 python scripts/bench_summary.py E1024 --steps 5 --warmup 3 --rows 100000000
MVF_B200_LIB=$PWD/metrovector_b200/libmvf_b200_stats.so timeout 300 python bench.py --workload E1024 --rows 100000000 --steps 3 --warmup 3 --no-cpu 2>&1 | grep "tc stats" | tail -1
timeout 400 python scripts/bench_summary.py E1024 --steps 5 --warmup 3
