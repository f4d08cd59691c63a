#!/bin/bash
# 8-GPU pass over the batched (tensor-path) configs only; no multirank check (run at 2 GPUs), each bench under its own timeout
for W in D1024 C E1024; do
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --workload $W --steps 6 --warmup 3 --no-cpu 2>&1 | grep '^{' | tee gpurun_out/bench_${W}_n8.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); r = d['roofline']
    print(d['config']['workload'], '| gpus', d['n_gpus'], '| qps', round(d['value'], 1), '| ms', round(d['ms_per_step'], 4), '| e2e', round(d['e2e']['value'], 1), '|', r['kernel'], round(r['kernel_ms'], 4), 'ms', round(r['achieved']), r['unit'], round(r['frac'], 3), '| clk', d['clocks']['sm_mhz'], d['clocks']['reasons'], '| ok', d['verified_against_oracle'], d.get('tensor_path_cross_checked_vs_scan'))
"
done
