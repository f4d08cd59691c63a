#!/bin/bash
# tests, then same-box A/B of two knobs on the long-row batched configs
timeout 200 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -3
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "tests failed: benches skipped"; exit 1; }
for rep in 1 2; do for V in "MVF_TC_PAIR=0" "MVF_TC_PAIR=1" "MVF_TC_PAIR=1 MVF_TC_BIGCAP=1"; do
echo "== $V (rep $rep)"
env $V timeout 300 python scripts/bench_summary.py C --steps 6 --warmup 3
env $V timeout 300 python scripts/bench_summary.py D1024 --steps 6 --warmup 3 --rows 20000000
done; done
