#!/bin/bash
# latency-bound shapes after a scan-kernel change: parity first, then A, the 125k-row shard of B, B, and D for reference
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_nonfinite.py tests/test_gpu_fullsize.py -m gpu -q -x --timeout 600 2>&1 | tail -3
for r in 256 100000; do timeout 120 python scripts/bench_summary.py A --steps 500 --warmup 50 --rows $r; done
timeout 300 python scripts/bench_summary.py B --steps 300 --warmup 30 --rows 125000 --no-batched
timeout 300 python scripts/bench_summary.py B --steps 100 --warmup 10 --no-batched
timeout 300 python scripts/bench_summary.py D --steps 20 --warmup 5 --no-batched
