#!/bin/bash
# usual GPU round: parity tests, then the bench on the main workloads
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python scripts/bench_summary.py B D E A --steps 30 --warmup 5
