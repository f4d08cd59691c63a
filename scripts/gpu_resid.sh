#!/bin/bash
for T in 2 4 8 12; do echo "threads $T"; MVF_UPLOAD_THREADS=$T timeout 600 python scripts/residency_bench.py | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print([(u['path'], round(u['gb_per_s'], 1), u['search_identical_to_generated_space']) for u in d['uploads']], d['uploads'][0]['note'])
"; done
