#!/bin/bash
# N GPUs: the peer-memory exchange (X1) in both process models against the single-GPU answer, then an A/B of the
# exchange route (peer kernel vs ncclAllGather + merge) on the latency-bound workloads
mkdir -p gpurun_out/peer
N=$(nvidia-smi -L | wc -l)
echo "gpus: $N"
timeout 300 python tests/multidevice_check.py > gpurun_out/peer/multidevice_n$N.txt 2>&1; echo "multidevice rc=$?"; tail -4 gpurun_out/peer/multidevice_n$N.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multirank_check.py > gpurun_out/peer/multirank_n$N.txt 2>&1; echo "multirank rc=$?"; tail -4 gpurun_out/peer/multirank_n$N.txt
for wl in B A D; do
  steps=300; [ $wl = D ] && steps=30
  for px in 1 0; do
    MVF_PEER_EXCHANGE=$px timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload $wl --steps $steps --warmup 20 --no-cpu --no-batched > gpurun_out/peer/bench_${wl}_n${N}_px$px.json 2> gpurun_out/peer/bench_${wl}_n${N}_px$px.err; echo "bench $wl px=$px rc=$?"
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/peer/bench_${wl}_n${N}_px$px.json").read().strip().splitlines()[-1])
    print("$wl px=$px", "ms/step", round(d["ms_per_step"], 5), "e2e ms", round(d["e2e"]["ms_per_step"], 5), "launches", d.get("gpu_launches"), "parity", d.get("multirank_parity"))
except Exception as e:
    print("no line", e)
PY
  done
done
