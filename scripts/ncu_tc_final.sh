#!/bin/bash
# Round-1 tensor-kernel profile set (after pacing + integer epilogue rework).  Every ncu run follows a plain run of
# the same command that exited 0.  .ncu-rep files exceed the return limit, so pages are exported on the box.
exp() {  # $1 = tag
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/prof_$1_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page source --csv > gpurun_out/prof_$1_source.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page details > gpurun_out/prof_$1_details.txt 2>/dev/null
}
rm -f gpurun_out/prof_*
C="python bench.py --workload C --steps 2 --warmup 3 --no-cpu"
$C > gpurun_out/plain_C.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_topk_tc -s 3 -c 1 -o /tmp/prof_C $C > gpurun_out/ncu_C.log 2>&1; exp C
$C > gpurun_out/plain_C.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_C.csv $C > gpurun_out/ncu_launch_C.log 2>&1
D="python bench.py --workload D1024 --rows 20000000 --steps 2 --warmup 3 --no-cpu"
$D > gpurun_out/plain_D1024.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_topk_tc -s 3 -c 1 -o /tmp/prof_D1024 $D > gpurun_out/ncu_D1024.log 2>&1; exp D1024
E1="python bench.py --workload E1024 --rows 100000000 --steps 2 --warmup 3 --no-cpu"
$E1 > gpurun_out/plain_E1024.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_topk_tc -s 3 -c 1 -o /tmp/prof_E1024 $E1 > gpurun_out/ncu_E1024.log 2>&1; exp E1024
ls -la gpurun_out | grep prof_
for f in C D1024 E1024; do tail -n 2 gpurun_out/ncu_$f.log; done
