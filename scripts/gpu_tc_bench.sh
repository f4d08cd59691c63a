#!/bin/bash
timeout 600 python scripts/bench_summary.py C --steps 5 --warmup 3
timeout 600 python scripts/bench_summary.py D1024 --steps 5 --warmup 3 --rows 20000000
timeout 600 python scripts/bench_summary.py E1024 --steps 5 --warmup 3 --rows 100000000
