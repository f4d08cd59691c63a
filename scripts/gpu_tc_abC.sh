#!/bin/bash
# same-box A/B of a tensor-path knob on config C only (tests first, short timeout)
VAR=${1:-MVF_TC_SYNC_WINDOW}; shift
timeout 200 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -15
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "tests failed: benches skipped"; exit 1; }
for rep in 1 2; do for V in "$@"; do
echo "== $VAR=$V (rep $rep)"
env $VAR=$V timeout 300 python scripts/bench_summary.py C --steps 6 --warmup 3
done; done
