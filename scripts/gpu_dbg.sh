#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_tc.py -x -q -k "paths_agree" 2>&1 | tail -40
