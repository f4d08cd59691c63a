#!/bin/bash
# same-box A/B of library builds kept under ab/ (not tracked): tests on the new build, then the batched configs on both
cp ab/lib_new.so metrovector_b200/libmvf_b200.so
timeout 250 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -3
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "tests failed: benches skipped"; exit 1; }
for rep in 1 2; do for L in head new; do
echo "== lib_$L (rep $rep)"
cp ab/lib_$L.so metrovector_b200/libmvf_b200.so
timeout 300 python scripts/bench_summary.py C --steps 6 --warmup 3
timeout 300 python scripts/bench_summary.py D1024 --steps 6 --warmup 3 --rows 20000000
done; done
