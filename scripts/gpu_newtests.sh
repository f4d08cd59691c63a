#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "crc32 or get_vectors" 2>&1 | tail -15
MVF_UPLOAD_THREADS=8 timeout 600 python scripts/residency_bench.py
