#!/bin/bash
# session check of HEAD on one GPU: the whole GPU test-suite (with durations), smoke(), the default bench line
mkdir -p gpurun_out/verify
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --durations=15 ) > gpurun_out/verify/pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/verify/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 600 python bench.py --steps 20 --warmup 5 ) > gpurun_out/verify/bench_default.json 2> gpurun_out/verify/bench_default.err; echo "bench rc=$?"; cat gpurun_out/verify/bench_default.json; tail -4 gpurun_out/verify/bench_default.err
