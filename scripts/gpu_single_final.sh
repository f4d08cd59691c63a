#!/bin/bash
WORKLOADS="D1024 E1024" ./scripts/gpu_multi.sh 1
