#!/bin/bash
timeout 250 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -3
timeout 200 python scripts/bench_summary.py C --steps 4 --warmup 3
