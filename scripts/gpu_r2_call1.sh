#!/bin/bash
# round 2, first GPU call: microbenchmark, the whole GPU test-suite, then the bench lines of every workload (1 GPU)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi.txt 2>&1
timeout 120 scripts/microbench/tmem_drain > gpurun_out/tmem_drain.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 --timeout-method thread > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -5 gpurun_out/pytest.log
for wl in D B C E; do
  steps=20; [ $wl = C ] && steps=5
  timeout 600 python bench.py --workload $wl --steps $steps --warmup 5 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
  echo "bench $wl rc=$?"
done
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "ref rc=$?"
