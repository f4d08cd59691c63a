#!/bin/bash
# tcgen05.cp as accumulator seeder: semantics (SBO = 0 broadcast) and cost
mkdir -p gpurun_out
timeout 60 ./scripts/microbench/bin/tmem_cp_seed 2>&1 | tee gpurun_out/tmem_cp_seed.txt
