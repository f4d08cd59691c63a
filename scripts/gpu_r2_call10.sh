#!/bin/bash
echo "== C default"; timeout 300 python scripts/bench_summary.py C --steps 5 --warmup 3
echo "== C EG=4"; MVF_TC_EPI_GROUPS=4 timeout 300 python scripts/bench_summary.py C --steps 5 --warmup 3
echo "== C EG=4 stats"; MVF_TC_EPI_GROUPS=4 MVF_B200_LIB=$PWD/metrovector_b200/libmvf_b200_stats.so timeout 300 python bench.py --workload C --steps 3 --warmup 3 --no-cpu 2>&1 | grep "tc stats" | tail -1
echo "== C items_per_sm=4"; MVF_TC_ITEMS_PER_SM=4 timeout 300 python scripts/bench_summary.py C --steps 5 --warmup 3
echo "== D1024 slab EG=4"; MVF_TC_EPI_GROUPS=4 timeout 300 python scripts/bench_summary.py D1024 --steps 5 --warmup 3 --rows 20000000
echo "== D1024 slab default"; timeout 300 python scripts/bench_summary.py D1024 --steps 5 --warmup 3 --rows 20000000
