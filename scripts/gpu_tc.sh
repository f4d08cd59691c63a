#!/bin/bash
# tensor-core path bring-up: run under a timeout so a barrier bug cannot hold the box
timeout 300 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -30
