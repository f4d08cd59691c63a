#!/bin/bash
# residency with the default staging-thread count (host cores - 4, capped at 12) against 8 threads, 3.07 GB file
nproc
for T in default 8; do
  if [ $T = default ]; then unset MVF_UPLOAD_THREADS; else export MVF_UPLOAD_THREADS=$T; fi
  echo "threads $T"; timeout 600 python scripts/residency_bench.py | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print([(u['variant'], u['path'][:14], round(u['gb_per_s'], 1), u['search_identical_to_generated_space']) for u in d['uploads']])
"; done
