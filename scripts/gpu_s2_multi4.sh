#!/bin/bash
# 4 GPUs: the two-process-model checks (the driver's scaling run includes N = 4)
N=$(nvidia-smi -L | wc -l)
echo "gpus: $N"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multirank_check.py 2>&1 | tail -3
