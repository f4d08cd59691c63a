#!/bin/bash
# what the driver runs at round end: GPU tests, smoke(), the default bench (with its CPU baseline), plus the small-shape kernel sweep
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python tests/sanitize_small.py 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 4 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 1800 gpurun_out/bench_default.json
