#!/bin/bash
# ncu captures for the scan kernel: launch list for B, full sets for B / D (10M-row slab) / E (100M-row slab).
# Reports are exported to CSV on the box (the .ncu-rep files exceed the 64 MiB return limit).
B="python bench.py --workload B --steps 3 --warmup 3 --no-cpu"
D="python bench.py --workload D --rows 10000000 --steps 3 --warmup 3 --no-cpu"
E="python bench.py --workload E --rows 100000000 --steps 3 --warmup 3 --no-cpu"
$B > gpurun_out/plain_B.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_B.csv $B > gpurun_out/ncu_launch_B.log 2>&1
for W in B D E; do
  eval CMD=\$$W
  $CMD > gpurun_out/plain_$W.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 4 -c 1 -o /tmp/prof_$W $CMD > gpurun_out/ncu_$W.log 2>&1
  ncu -i /tmp/prof_$W.ncu-rep --page raw --csv > gpurun_out/prof_${W}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$W.ncu-rep --page source --csv > gpurun_out/prof_${W}_source.csv 2>/dev/null
  ncu -i /tmp/prof_$W.ncu-rep --page details > gpurun_out/prof_${W}_details.txt 2>/dev/null
done
ls -la gpurun_out
