#!/bin/bash
# 2-GPU box: the whole GPU suite (the multi-GPU tests run instead of skipping)
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 ) 2>&1 | tail -8
