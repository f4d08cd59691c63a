#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 --timeout-method thread > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -4 gpurun_out/pytest.log
for wl in E1024 D1024 C B1024; do
  timeout 600 python scripts/bench_summary.py $wl --steps 5 --warmup 3
done
