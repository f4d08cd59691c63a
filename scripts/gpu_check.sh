#!/bin/bash
# usual GPU round: parity tests, then the bench on the main workloads
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python scripts/bench_summary.py B D E A --steps 30 --warmup 5
timeout 600 python scripts/bench_summary.py C --steps 8 --warmup 3
timeout 600 python scripts/bench_summary.py D1024 --steps 8 --warmup 3 --rows 20000000
timeout 600 python scripts/bench_summary.py E1024 --steps 8 --warmup 3 --rows 100000000
