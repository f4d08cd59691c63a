#!/bin/bash
for D in 3 2; do echo "== dtype $D"; timeout 200 python scripts/dbg_tc.py $D 2>&1 | tail -4; done
timeout 300 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -3
