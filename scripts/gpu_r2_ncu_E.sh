#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --workload E1024 --rows 100000000 --steps 3 --warmup 3 --no-cpu"
$CMD > gpurun_out/ncu_E_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_topk_tc -s 2 -c 1 -o gpurun_out/r2_E1024 -f $CMD > gpurun_out/ncu_E.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_E.log
