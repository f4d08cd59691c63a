#!/bin/bash
# same-box A/B of a tensor-path knob on the two long-row batched configs, no tests
VAR=${1:-MVF_TC_PAIR}; shift
for rep in 1 2; do for V in "$@"; do
echo "== $VAR=$V (rep $rep)"
env $VAR=$V timeout 300 python scripts/bench_summary.py C --steps 6 --warmup 3
env $VAR=$V timeout 300 python scripts/bench_summary.py D1024 --steps 6 --warmup 3 --rows 20000000
done; done
