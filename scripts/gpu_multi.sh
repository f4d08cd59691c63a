#!/bin/bash
# multi-GPU: sharded search == single GPU == oracle, then the bench at N ranks
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multirank_check.py 2>&1 | tail -15
for W in ${WORKLOADS:-B D}; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload $W --steps ${STEPS:-30} --warmup 5 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); r = d['roofline']
    print(d['config']['workload'], 'gpus', d['n_gpus'], 'qps', round(d['value'], 1), 'ms', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1), r['kernel'], round(r['kernel_ms'], 4), round(r['achieved']), r['unit'], round(r['frac'], 3), d['clocks'], d['verified_against_oracle'])
"
done
