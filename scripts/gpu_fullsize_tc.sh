#!/bin/bash
# full-size single-GPU runs of the two batched configs that are benchmarked on slabs elsewhere
timeout 200 python scripts/bench_summary.py D1024 --steps 4 --warmup 3
cp gpurun_out/bench_D1024.json gpurun_out/bench_D1024_full_n1.json
timeout 200 python scripts/bench_summary.py E1024 --steps 4 --warmup 3
cp gpurun_out/bench_E1024.json gpurun_out/bench_E1024_full_n1.json
