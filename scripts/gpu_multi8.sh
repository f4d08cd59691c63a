#!/bin/bash
WORKLOADS="D D1024 E E1024 C B" ./scripts/gpu_multi.sh 8
