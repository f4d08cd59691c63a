#!/bin/bash
# ncu capture of the tensor-core kernel on a 4M-row slab of config D (batch 1024); CSV exported on the box
CMD="python bench.py --workload D1024 --rows 20000000 --steps 2 --warmup 3 --no-cpu"
$CMD > gpurun_out/plain_tc.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_topk_tc -s 3 -c 1 -o /tmp/prof_tc $CMD > gpurun_out/ncu_tc.log 2>&1
ncu -i /tmp/prof_tc.ncu-rep --page raw --csv > gpurun_out/prof_tc_raw.csv 2>/dev/null
ncu -i /tmp/prof_tc.ncu-rep --page source --csv > gpurun_out/prof_tc_source.csv 2>/dev/null
ncu -i /tmp/prof_tc.ncu-rep --page details > gpurun_out/prof_tc_details.txt 2>/dev/null
tail -3 gpurun_out/ncu_tc.log
