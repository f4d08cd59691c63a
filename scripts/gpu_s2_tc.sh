#!/bin/bash
# tensor-path tests (k up to 512) and the batched bench lines that must not have moved
mkdir -p gpurun_out/tc
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -x --timeout 600 2>&1 | tail -8
timeout 300 python scripts/bench_summary.py C --steps 5 --warmup 3
