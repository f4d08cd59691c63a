#!/bin/bash
# source-level ncu capture of the scan kernel on config A (latency-bound, 100k x 128 f32)
mkdir -p gpurun_out/ncuA
A="python bench.py --workload A --steps 5 --warmup 3 --no-cpu"
$A > gpurun_out/ncuA/plain_A.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 4 -c 1 -o /tmp/prof_A $A > gpurun_out/ncuA/ncu_A.log 2>&1
ncu -i /tmp/prof_A.ncu-rep --page raw --csv > gpurun_out/ncuA/prof_A_raw.csv 2>/dev/null
ncu -i /tmp/prof_A.ncu-rep --page source --csv > gpurun_out/ncuA/prof_A_source.csv 2>/dev/null
ncu -i /tmp/prof_A.ncu-rep --page source --csv --print-source cuda > gpurun_out/ncuA/prof_A_cuda.csv 2>/dev/null
ncu -i /tmp/prof_A.ncu-rep --page details > gpurun_out/ncuA/prof_A_details.txt 2>/dev/null
ls -la gpurun_out/ncuA
python scripts/bench_summary.py A --steps 300 --warmup 30
