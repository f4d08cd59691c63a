#!/bin/bash
# 2 GPUs: whole GPU suite (incl. the 2-GPU tests) after the local-id / ranked-merge change
N=$(nvidia-smi -L | wc -l)
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -8
timeout 300 python tests/multidevice_check.py 2>&1 | tail -4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multirank_check.py 2>&1 | tail -4
MVF_PEER_EXCHANGE=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tests/multirank_check.py 2>&1 | tail -3
