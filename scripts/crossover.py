#!/usr/bin/env python
"""Batch-size crossover between the scan kernel (K1, <= 8 queries per pass) and the tensor-core kernel (K2)
on a 20M x 768 f16 cosine slab and a 100M x 128 u8 L2 slab: ms per batch for each path."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import metrovector_b200 as mv  # noqa: E402

eng = mv.GpuEngine(0)
stream = torch.cuda.Stream()
eng.set_stream(stream.cuda_stream)
for name, dt, metric, n, d in (("f16 cos 20Mx768", 1, 2, 20_000_000, 768), ("u8 L2 100Mx128", 3, 0, 100_000_000, 128)):
    gs = eng.synthetic(dt, metric, n, d, 1)
    for nq in (4, 8, 12, 16, 32, 64, 128, 256):
        q = torch.from_numpy(mv.synth_fill_host(dt, 2, 0, nq, d)).cuda()
        ids = torch.empty((nq, 10), dtype=torch.int64, device="cuda")
        sc = torch.empty((nq, 10), dtype=torch.float32, device="cuda")
        row = []
        for path in (1, 2):
            eng.set_path(path)
            with torch.cuda.stream(stream):
                for _ in range(2):
                    gs.search_device(q.data_ptr(), dt, nq, 10, metric, ids.data_ptr(), sc.data_ptr())
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                for _ in range(3):
                    gs.search_device(q.data_ptr(), dt, nq, 10, metric, ids.data_ptr(), sc.data_ptr())
                b.record(stream)
                torch.cuda.synchronize()
            row.append(a.elapsed_time(b) / 3)
        print("%-18s nq %4d   scan %9.3f ms   tensor %9.3f ms   -> %s" % (name, nq, row[0], row[1], "tensor" if row[1] < row[0] else "scan"), flush=True)
    gs.free()
eng.set_path(0)
