#!/bin/bash
# Round-2 profile set.  Every ncu run follows a plain run of the same command that exited 0.  .ncu-rep files exceed the
# return limit, so pages are exported on the box (raw + source + details) and condensed here by profiles/summarize.py.
exp() {  # $1 = tag
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/prof_$1_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page source --csv > gpurun_out/prof_$1_source.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page details > gpurun_out/prof_$1_details.txt 2>/dev/null
}
full() {  # $1 = tag, $2 = kernel regex, rest = command
  tag=$1; rx=$2; shift 2
  "$@" > gpurun_out/plain_$tag.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 -f -o /tmp/prof_$tag "$@" > gpurun_out/ncu_$tag.log 2>&1
  exp $tag; tail -n 1 gpurun_out/ncu_$tag.log
}
mkdir -p gpurun_out; rm -f gpurun_out/prof_* gpurun_out/launches_*
DEF="python bench.py --steps 2 --warmup 3 --no-cpu"
# the default bench command: launch list (both legs), then its two dominant kernels
$DEF > gpurun_out/plain_default.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_default.csv $DEF > gpurun_out/ncu_launch_default.log 2>&1
full D scan_topk $DEF
full D1024 gemm_topk_tc $DEF
full C gemm_topk_tc python bench.py --workload C --steps 2 --warmup 3 --no-cpu
full E1024 gemm_topk_tc python bench.py --workload E1024 --steps 2 --warmup 3 --no-cpu
full B1024 gemm_topk_tc python bench.py --workload B1024 --steps 2 --warmup 3 --no-cpu
full B scan_topk python bench.py --workload B --steps 2 --warmup 3 --no-cpu --no-batched
full A scan_topk python bench.py --workload A --steps 2 --warmup 3 --no-cpu
ls -la gpurun_out | grep prof_ | head -30
