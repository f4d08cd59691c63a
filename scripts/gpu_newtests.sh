#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "threads or external_ids or profile" 2>&1 | tail -6
timeout 200 python scripts/bench_summary.py A B --steps 30 --warmup 5
