#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "external_ids or tombstone or row_mask" 2>&1 | tail -15
