#!/bin/bash
WORKLOADS="B D" ./scripts/gpu_multi.sh 8
STEPS=5 WORKLOADS="D1024" ./scripts/gpu_multi.sh 8 | tail -2
