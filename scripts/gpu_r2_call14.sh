#!/bin/bash
for c in 0 1 2 4; do echo "== A ctas_per_sm=$c"; MVF_SCAN_CTAS_PER_SM=$c timeout 300 python scripts/bench_summary.py A --steps 300 --warmup 30; done
for c in 0 1 2; do echo "== B ctas_per_sm=$c"; MVF_SCAN_CTAS_PER_SM=$c timeout 300 python scripts/bench_summary.py B --steps 100 --warmup 10 --no-batched; done
echo "== B 8-GPU shard (125k rows)"; timeout 300 python scripts/bench_summary.py B --steps 300 --warmup 30 --rows 125000 --no-batched
MVF_SCAN_CTAS_PER_SM=1 timeout 300 python scripts/bench_summary.py B --steps 300 --warmup 30 --rows 125000 --no-batched
