#!/bin/bash
# tests, then same-box A/B of one tensor-path knob on the three batched configs ("auto" = knob unset)
VAR=${1:-MVF_TC_SYNC_WINDOW}; shift
timeout 250 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -3
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "tests failed: benches skipped"; exit 1; }
for rep in 1 2; do for V in "$@"; do
echo "== $VAR=$V (rep $rep)"
if [ "$V" = auto ]; then E=""; else E="$VAR=$V"; fi
env $E timeout 300 python scripts/bench_summary.py C --steps 6 --warmup 3
env $E timeout 300 python scripts/bench_summary.py D1024 --steps 6 --warmup 3 --rows 20000000
env $E timeout 300 python scripts/bench_summary.py E1024 --steps 6 --warmup 3 --rows 100000000
done; done
