#!/bin/bash
# diagnostic build (-DMVF_TC_STATS): per-role cycle breakdown of the tensor kernel on the batched configs
for W in "C" "D1024 --rows 20000000" "E1024 --rows 100000000"; do
echo "== $W"
timeout 300 python bench.py --workload $W --steps 3 --warmup 2 --no-cpu 2>&1 | grep "tc stats" | tail -3
done
echo "== C multicast"; MVF_TC_PAIR=0 timeout 300 python bench.py --workload C --steps 3 --warmup 2 --no-cpu 2>&1 | grep "tc stats" | tail -2
echo "== C no pacing"; MVF_TC_SYNC_WINDOW=0 timeout 300 python bench.py --workload C --steps 3 --warmup 2 --no-cpu 2>&1 | grep "tc stats" | tail -2
