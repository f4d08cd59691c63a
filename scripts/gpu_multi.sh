#!/bin/bash
# multi-GPU: sharded search == single GPU == oracle, then the bench at N ranks for the given workloads
N=${1:-2}
if [ "$N" -gt 1 ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multirank_check.py 2>&1 | grep -E "multirank|MISMATCH|Error|error" | head
fi
for W in ${WORKLOADS:-D}; do
case $W in C|D1024|E1024) STEPS=6;; *) STEPS=30;; esac
if [ "$N" -gt 1 ]; then LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"; else LAUNCH="python"; fi
timeout 900 $LAUNCH bench.py --gpus $N --workload $W --steps $STEPS --warmup 4 --no-cpu 2>&1 | grep '^{' | tee gpurun_out/bench_${W}_n${N}.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); r = d['roofline']
    print(d['config']['workload'], '| gpus', d['n_gpus'], '| qps', round(d['value'], 1), '| ms', round(d['ms_per_step'], 4), '| e2e', round(d['e2e']['value'], 1), '|', r['kernel'], round(r['kernel_ms'], 4), 'ms', round(r['achieved']), r['unit'], round(r['frac'], 3), '| clk', d['clocks']['sm_mhz'], d['clocks']['reasons'], '| ok', d['verified_against_oracle'])
"
done
