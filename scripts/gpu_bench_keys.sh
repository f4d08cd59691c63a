#!/bin/bash
# quick check that both bench branches (scan, tensor) still emit their JSON line after an edit to bench.py
timeout 100 python bench.py --workload A --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('A', d['roofline']['frac'], d['roofline']['frac_of_nominal'])"
timeout 100 python bench.py --workload C --rows 300000 --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('C', r['frac'], r['frac_of_nominal'], r['peak_source'], d['tensor_path_cross_checked_vs_scan'])"
