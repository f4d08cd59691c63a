#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_tc.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -q --timeout 600 --timeout-method thread > gpurun_out/pytest2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest2.log
tail -8 gpurun_out/pytest2.log
timeout 600 python bench.py --workload E --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_E2.json 2> gpurun_out/bench_E2.err
echo "bench E rc=$?"
