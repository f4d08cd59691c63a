#!/bin/bash
# Round-1 profile set.  Every ncu run follows a plain run of the same command that exited 0.
# .ncu-rep files exceed the 64 MiB return limit, so pages are exported to CSV/text on the box.
exp() {  # $1 = tag
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/prof_$1_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page source --csv > gpurun_out/prof_$1_source.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page details > gpurun_out/prof_$1_details.txt 2>/dev/null
}
DEF="python bench.py --steps 5 --warmup 3 --no-cpu"                       # the default bench command (config D, full size)
$DEF > gpurun_out/plain_default.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_default.csv $DEF > gpurun_out/ncu_launch_default.log 2>&1
$DEF > gpurun_out/plain_default.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 4 -c 1 -o /tmp/prof_D $DEF > gpurun_out/ncu_D.log 2>&1; exp D
B="python bench.py --workload B --steps 5 --warmup 3 --no-cpu"
$B > gpurun_out/plain_B.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 4 -c 1 -o /tmp/prof_B $B > gpurun_out/ncu_B.log 2>&1; exp B
E="python bench.py --workload E --rows 200000000 --steps 3 --warmup 3 --no-cpu"
$E > gpurun_out/plain_E.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 4 -c 1 -o /tmp/prof_E $E > gpurun_out/ncu_E.log 2>&1; exp E
C="python bench.py --workload C --steps 2 --warmup 3 --no-cpu"
$C > gpurun_out/plain_C.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_topk_tc -s 3 -c 1 -o /tmp/prof_C $C > gpurun_out/ncu_C.log 2>&1; exp C
$C > gpurun_out/plain_C.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_C.csv $C > gpurun_out/ncu_launch_C.log 2>&1
ls -la gpurun_out | head -40
E1="python bench.py --workload E1024 --rows 100000000 --steps 2 --warmup 3 --no-cpu"
$E1 > gpurun_out/plain_E1024.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_topk_tc -s 3 -c 1 -o /tmp/prof_E1024 $E1 > gpurun_out/ncu_E1024.log 2>&1; exp E1024
ls gpurun_out | head -50
