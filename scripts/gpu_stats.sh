#!/bin/bash
# diagnostic build (-DMVF_TC_STATS, metrovector_b200/libmvf_b200_stats.so): per-role cycle breakdown of the tensor kernel
export MVF_B200_LIB=$PWD/metrovector_b200/libmvf_b200_stats.so
for W in "E1024 --rows 100000000" "D1024 --rows 20000000" "C" "B1024"; do
echo "== $W"
timeout 300 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu 2>&1 | grep "tc stats" | tail -2
done
