#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -x --timeout 600 -k randomised 2>&1 | tail -25
