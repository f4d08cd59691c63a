#!/bin/bash
# round 2, single-GPU closing run: the whole GPU test-suite, smoke(), every workload's bench line, the reference arm
mkdir -p gpurun_out/final
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/final/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/final/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for wl in D B C E A; do
  steps=20; [ $wl = C ] && steps=5; [ $wl = A ] && steps=200
  timeout 600 python bench.py --workload $wl --steps $steps --warmup 5 > gpurun_out/final/bench_${wl}_n1.json 2> gpurun_out/final/bench_${wl}_n1.err; echo "bench $wl rc=$?"
done
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/final/bench_ref_n1.json 2> gpurun_out/final/bench_ref.err; echo "ref rc=$?"
