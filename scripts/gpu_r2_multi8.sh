#!/bin/bash
# N GPUs of one box: both process models against the single-GPU answer and the oracle, then the bench lines
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "gpus: $N"
timeout 600 python tests/multidevice_check.py > gpurun_out/multidevice_n$N.txt 2>&1; echo "multidevice rc=$?"; tail -2 gpurun_out/multidevice_n$N.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multirank_check.py > gpurun_out/multirank_n$N.txt 2>&1; echo "multirank rc=$?"; tail -2 gpurun_out/multirank_n$N.txt
for wl in D E C B; do
  steps=20; [ $wl = C ] && steps=5
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload $wl --steps $steps --warmup 5 > gpurun_out/bench_${wl}_n$N.json 2> gpurun_out/bench_${wl}_n$N.err; echo "bench $wl rc=$?"
done
