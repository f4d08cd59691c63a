#!/bin/bash
for r in 256 4096 20000 50000 100000 200000 400000; do MVF_NO_FUSED_MERGE=1 timeout 120 python scripts/bench_summary.py A --steps 300 --warmup 30 --rows $r 2>&1 | sed "s/^/sep   /"; done
for r in 256 4096 20000 50000 100000 200000 400000; do timeout 120 python scripts/bench_summary.py A --steps 300 --warmup 30 --rows $r 2>&1 | sed "s/^/fused /"; done
