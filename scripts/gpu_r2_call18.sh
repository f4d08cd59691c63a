#!/bin/bash
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_nonfinite.py -m gpu -q -x --timeout 900 2>&1 | tail -4
for r in 256 100000; do timeout 120 python scripts/bench_summary.py A --steps 300 --warmup 30 --rows $r; done
timeout 300 python scripts/bench_summary.py B --steps 100 --warmup 10 --no-batched
timeout 300 python scripts/bench_summary.py B --steps 300 --warmup 30 --rows 125000 --no-batched
timeout 300 python scripts/bench_summary.py D --steps 20 --warmup 5 --no-batched
timeout 300 python scripts/bench_summary.py E --steps 20 --warmup 5 --no-batched
