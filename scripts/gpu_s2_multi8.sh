#!/bin/bash
# 8 GPUs of one box: both process models against the single-GPU answer and the oracle (logs kept), then bench lines
mkdir -p gpurun_out/n8
N=$(nvidia-smi -L | wc -l)
echo "gpus: $N"
timeout 400 python tests/multidevice_check.py > gpurun_out/n8/multidevice_n$N.txt 2>&1; echo "multidevice rc=$?"; tail -2 gpurun_out/n8/multidevice_n$N.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multirank_check.py > gpurun_out/n8/multirank_n$N.txt 2>&1; echo "multirank rc=$?"; tail -2 gpurun_out/n8/multirank_n$N.txt
for wl in B A D C; do
  steps=20; extra=""
  [ $wl = C ] && steps=5
  [ $wl = B ] && steps=300
  [ $wl = A ] && steps=500
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload $wl --steps $steps --warmup 10 --no-cpu > gpurun_out/n8/bench_${wl}_n$N.json 2> gpurun_out/n8/bench_${wl}_n$N.err; echo "bench $wl rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/n8/bench_${wl}_n$N.json").read().strip().splitlines()[-1])
    b = d.get("batched") or {}
    print("$wl", "qps", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 5), "e2e ms", round(d["e2e"]["ms_per_step"], 5), "launches", d.get("gpu_launches"), "parity", d.get("multirank_parity"), "| batched", b.get("value"), b.get("ms_per_step"), (b.get("roofline") or {}).get("frac"))
except Exception as e:
    print("no line", e)
PY
done
MVF_PEER_EXCHANGE=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload B --steps 300 --warmup 10 --no-cpu --no-batched > gpurun_out/n8/bench_B_n${N}_nccl.json 2> gpurun_out/n8/bench_B_n${N}_nccl.err; echo "bench B nccl rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/n8/bench_B_n${N}_nccl.json").read().strip().splitlines()[-1])
print("B (NCCL exchange)", "qps", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 5), "e2e ms", round(d["e2e"]["ms_per_step"], 5))
PY
