#!/bin/bash
# tcgen05.cp as accumulator seeder: semantics (SBO = 0 broadcast) and cost
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tmem_cp_seed scripts/microbench/tmem_cp_seed.cu && timeout 60 /tmp/tmem_cp_seed 2>&1 | tee gpurun_out/tmem_cp_seed.txt
