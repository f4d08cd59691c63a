#!/bin/bash
CMD="python bench.py --workload E1024 --rows 20000000 --steps 2 --warmup 3 --no-cpu"
$CMD > gpurun_out/plain_E1024.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_topk_tc -s 3 -c 1 -o /tmp/prof_E1024 $CMD > gpurun_out/ncu_E1024.log 2>&1
ncu -i /tmp/prof_E1024.ncu-rep --page raw --csv > gpurun_out/prof_E1024_raw.csv 2>/dev/null
ncu -i /tmp/prof_E1024.ncu-rep --page source --csv > gpurun_out/prof_E1024_source.csv 2>/dev/null
ncu -i /tmp/prof_E1024.ncu-rep --page details > gpurun_out/prof_E1024_details.txt 2>/dev/null
