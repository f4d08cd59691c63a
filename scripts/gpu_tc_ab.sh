#!/bin/bash
# same-box A/B of a tensor-path knob: alternates the settings so power/thermal state is shared.
# The tests run first under a short timeout; a failure (or a hang) skips the benches.
VAR=${1:-MVF_TC_EPI_GROUPS}; shift
timeout 150 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -15
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "tests failed: benches skipped"; exit 1; }
for rep in 1 2; do for V in "$@"; do
echo "== $VAR=$V (rep $rep)"
env $VAR=$V timeout 300 python scripts/bench_summary.py C --steps 6 --warmup 3
env $VAR=$V timeout 300 python scripts/bench_summary.py D1024 --steps 6 --warmup 3 --rows 20000000
env $VAR=$V timeout 300 python scripts/bench_summary.py E1024 --steps 6 --warmup 3 --rows 100000000
done; done
