#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_stats.sh > gpurun_out/tc_role_stats.txt 2>&1; cat gpurun_out/tc_role_stats.txt
for R in 1000000 3000000; do
ROWS=$R timeout 900 python scripts/residency_bench.py > gpurun_out/residency_$R.json 2> gpurun_out/residency_$R.err; echo "residency $R rc=$?"
python - gpurun_out/residency_$R.json <<'PY'
import sys, json
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("file %.2f GB" % d["file_gb"])
for u in d["uploads"]:
    print("%-8s %-42s %6.1f GB/s  identical=%s  note: %s" % (u["variant"], u["path"], u["gb_per_s"], u["search_identical_to_generated_space"], u["note"]))
PY
done
