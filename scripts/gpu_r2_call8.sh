#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x --timeout 600 --timeout-method thread -k "not tensor" 2>&1 | tail -3
for wl in A B; do
echo "== fused"; timeout 300 python scripts/bench_summary.py $wl --steps 200 --warmup 20 --no-batched
echo "== separate K3"; MVF_NO_FUSED_MERGE=1 timeout 300 python scripts/bench_summary.py $wl --steps 200 --warmup 20 --no-batched
done
echo "== D fused"; timeout 300 python scripts/bench_summary.py D --steps 20 --warmup 5 --no-batched
