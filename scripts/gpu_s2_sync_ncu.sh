#!/bin/bash
# (1) result wait: polling (default) vs MVF_SYNC_MODE=block on the e2e figures; (2) ncu refresh of the scan kernel (D, B, A)
# and the launch list of the default bench after this session's K1 changes.  Every ncu run follows a plain run that exited 0.
mkdir -p gpurun_out/s2
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 2>&1 | tail -2
for mode in poll block; do
  for wl in D B A; do
    steps=20; [ $wl = B ] && steps=200; [ $wl = A ] && steps=500
    MVF_SYNC_MODE=$mode timeout 300 python bench.py --workload $wl --steps $steps --warmup 5 --no-cpu --no-batched > gpurun_out/s2/bench_${wl}_$mode.json 2>/dev/null
    python - <<PY
import json
d = json.loads(open("gpurun_out/s2/bench_${wl}_$mode.json").read().strip().splitlines()[-1])
print("$wl $mode", "device ms", round(d["ms_per_step"], 5), "e2e ms", round(d["e2e"]["ms_per_step"], 5), "clk", d["clocks"]["sm_mhz"])
PY
  done
done
exp() {
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/s2/prof_$1_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page source --csv > gpurun_out/s2/prof_$1_source.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page details > gpurun_out/s2/prof_$1_details.txt 2>/dev/null
}
full() {
  tag=$1; rx=$2; shift 2
  "$@" > gpurun_out/s2/plain_$tag.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 -f -o /tmp/prof_$tag "$@" > gpurun_out/s2/ncu_$tag.log 2>&1
  exp $tag; tail -n 1 gpurun_out/s2/ncu_$tag.log
}
DEF="python bench.py --steps 2 --warmup 3 --no-cpu"
$DEF > gpurun_out/s2/plain_default.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/s2/launches_default.csv $DEF > gpurun_out/s2/ncu_launch_default.log 2>&1
full D scan_topk python bench.py --steps 2 --warmup 3 --no-cpu --no-batched
full B scan_topk python bench.py --workload B --steps 2 --warmup 3 --no-cpu --no-batched
full A scan_topk python bench.py --workload A --steps 2 --warmup 3 --no-cpu
rm -f gpurun_out/s2/prof_D_source.csv gpurun_out/s2/prof_B_source.csv.bak
ls -la gpurun_out/s2 | head -40
