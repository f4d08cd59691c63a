#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x --timeout 600 -k "not tensor" 2>&1 | tail -40
