#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -4
timeout 900 python scripts/crossover.py 2>/dev/null
