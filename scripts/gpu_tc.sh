#!/bin/bash
# tensor-core path: tests under a timeout (a barrier bug must not hold the box), then the batched benches
timeout 300 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -5
for dt in 3 2; do timeout 300 python scripts/tc_vs_scan.py $dt 20000000 128 256 10 2>&1 | tail -3; done
timeout 600 python scripts/bench_summary.py C --steps 8 --warmup 3
timeout 600 python scripts/bench_summary.py D1024 --steps 8 --warmup 3 --rows 20000000
timeout 600 python scripts/bench_summary.py E1024 --steps 8 --warmup 3 --rows 100000000
