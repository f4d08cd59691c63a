#!/bin/bash
# 2 GPUs: the one-process multi-device engine, the two-rank NCCL path, and a 2-rank bench line
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "gpus: $N"
timeout 600 python tests/multidevice_check.py > gpurun_out/multidevice_n$N.txt 2>&1; echo "multidevice rc=$?"; tail -5 gpurun_out/multidevice_n$N.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multirank_check.py > gpurun_out/multirank_n$N.txt 2>&1; echo "multirank rc=$?"; tail -5 gpurun_out/multirank_n$N.txt
timeout 600 python -m pytest tests/test_gpu_multidevice.py tests/test_gpu_multirank.py -m gpu -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_D_n$N.json 2> gpurun_out/bench_D_n$N.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_D_n$N.err
