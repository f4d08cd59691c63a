#!/bin/bash
mkdir -p gpurun_out
for eg in 4 2; do
echo "== E1024 slab EG=$eg (product build)"
MVF_TC_EPI_GROUPS=$eg timeout 300 python scripts/bench_summary.py E1024 --steps 5 --warmup 3 --rows 100000000
echo "== stats EG=$eg"
MVF_TC_EPI_GROUPS=$eg MVF_B200_LIB=$PWD/metrovector_b200/libmvf_b200_stats.so timeout 300 python bench.py --workload E1024 --rows 100000000 --steps 3 --warmup 3 --no-cpu 2>&1 | grep "tc stats" | tail -1
done
