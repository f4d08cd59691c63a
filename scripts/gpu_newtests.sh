#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py -x -q -k "large_k or tensor_path_search" 2>&1 | tail -15
