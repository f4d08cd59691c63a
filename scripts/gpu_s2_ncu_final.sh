#!/bin/bash
# final r2b profile set of the scan kernel (D, B, A) and the default launch list, after the last K1 changes.
# Every ncu run follows a plain run of the same command that exited 0.
mkdir -p gpurun_out/s2f
exp() {
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/s2f/prof_$1_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page source --csv > gpurun_out/s2f/prof_$1_source.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page details > gpurun_out/s2f/prof_$1_details.txt 2>/dev/null
}
full() {
  tag=$1; rx=$2; shift 2
  "$@" > gpurun_out/s2f/plain_$tag.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 -f -o /tmp/prof_$tag "$@" > gpurun_out/s2f/ncu_$tag.log 2>&1
  exp $tag; tail -n 1 gpurun_out/s2f/ncu_$tag.log
}
DEF="python bench.py --steps 2 --warmup 3 --no-cpu"
$DEF > gpurun_out/s2f/plain_default.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/s2f/launches_default.csv $DEF > gpurun_out/s2f/ncu_launch_default.log 2>&1
full D scan_topk python bench.py --steps 2 --warmup 3 --no-cpu --no-batched
full B scan_topk python bench.py --workload B --steps 2 --warmup 3 --no-cpu --no-batched
full A scan_topk python bench.py --workload A --steps 2 --warmup 3 --no-cpu
python - <<'PY'
# the source page of the 22 ms kernel is large: keep only what profiles/summarize.py reads (opcode + executed + stalls)
import csv, sys
csv.field_size_limit(1 << 30)
PY
ls -la gpurun_out/s2f | awk '{print $5, $9}' | tail -20
