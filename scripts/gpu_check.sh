#!/bin/bash
# usual GPU round: parity tests, then the bench on the main workloads
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python scripts/bench_summary.py B D E A --steps 30 --warmup 5
timeout 600 python scripts/bench_summary.py C --steps 8 --warmup 3
timeout 600 python scripts/bench_summary.py E1024 --steps 8 --warmup 3 --rows 100000000
timeout 600 python scripts/residency_bench.py
g++ -std=c++17 -Iinclude examples/similarity_search.cpp -Lmetrovector_b200 -lmvf_b200 -Wl,-rpath,$PWD/metrovector_b200 -o /tmp/similarity_search && /tmp/similarity_search | tail -22
