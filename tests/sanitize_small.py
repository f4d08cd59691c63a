"""Tiny end-to-end pass over every kernel (K0 norms, K1 wide/narrow scans, K3 merges, K2 tensor path,
synthetic fill), sized for `compute-sanitizer --tool memcheck`.  Exits non-zero on any mismatch."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import metrovector_b200 as mv  # noqa: E402
import oracle  # noqa: E402
from util import as_query, check_batch  # noqa: E402

eng = mv.GpuEngine(0)
for dtype, d, n in ((0, 768, 300), (0, 4, 60), (1, 100, 257), (2, 64, 500), (3, 128, 700), (3, 13, 90), (1, 768, 129)):
    db = mv.synth_fill_host(dtype, 5, 0, n, d)
    gs = eng.from_host(db, mv.DataType(dtype), mv.DistanceMetric.L2)
    for nq, k in ((1, 10), (5, 3), (9, 40)):
        q = np.stack([as_query(dtype, r) for r in mv.synth_fill_host(dtype, 6, 0, nq, d)])
        for metric in (0, 1, 2):
            ids, sc, isc = gs.search(q, k, metric=metric)
            check_batch(db, dtype, metric, q, k, ids, sc, isc)
    gs.free()
g2 = eng.synthetic(mv.DataType.UInt8, mv.DistanceMetric.L2, 3000, 128, seed=9)
assert np.array_equal(g2.read_rows(0, 3000), mv.synth_fill_host(3, 9, 0, 3000, 128))
eng.set_path(2)
for dtype, d, n in ((1, 128, 1500), (3, 128, 2000), (2, 96, 900)):
    db = mv.synth_fill_host(dtype, 7, 0, n, d)
    gs = eng.from_host(db, mv.DataType(dtype), mv.DistanceMetric.L2)
    q = mv.synth_fill_host(dtype, 8, 0, 140, d)
    for metric in ((0, 1, 2) if dtype == 1 else (0, 1)):
        ids, sc, isc = gs.search(q, 10, metric=metric)
        qq = q.astype(np.float32) if dtype == 1 else q
        check_batch(db, dtype, metric, qq[:4], 10, ids[:4], sc[:4], isc[:4])
        check_batch(db, dtype, metric, qq[-2:], 10, ids[-2:], sc[-2:], isc[-2:])
    gs.free()
eng.shutdown()
print("sanitize_small ok")
