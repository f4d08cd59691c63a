#!/bin/bash
VAR=${1}; A=${2}; B=${3}
for rep in 1 2; do for V in $A $B; do
echo "== $VAR=$V (rep $rep)"
env $VAR=$V timeout 600 python scripts/bench_summary.py C --steps 6 --warmup 3
done; done
