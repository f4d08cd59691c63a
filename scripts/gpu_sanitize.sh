#!/bin/bash
# one sanitizer tool per call (B200_PROFILING.md); the plain run must pass first
timeout 300 python tests/sanitize_small.py > gpurun_out/sanitize_plain.log 2>&1 && \
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python tests/sanitize_small.py > gpurun_out/sanitize_memcheck.log 2>&1
echo "rc=$?"
tail -3 gpurun_out/sanitize_plain.log; tail -8 gpurun_out/sanitize_memcheck.log
