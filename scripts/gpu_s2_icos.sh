#!/bin/bash
# integer cosine on the tensor path (MODE 3): parity, then the batched figure against the scan kernel's on a 100M-row u8 slab
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -q -x --timeout 600 2>&1 | tail -4
python - <<'PY'
import time, numpy as np, torch
import metrovector_b200 as mv
eng = mv.GpuEngine(0)
n, d, nq, k = 100_000_000, 128, 1024, 10
gs = eng.synthetic(mv.DataType.UInt8, mv.DistanceMetric.Cosine, n, d, 0x4D564631)
q = mv.synth_fill_host(3, 0x4D564632, 0, nq, d)
res = {}
for path in (2, 1):
    eng.set_path(path)
    m = nq if path == 2 else 64                      # the scan kernel takes 8 queries per pass: time 64 and scale
    gs.search(q[:m], k)
    t0 = time.perf_counter(); out = gs.search(q[:m], k); dt = time.perf_counter() - t0
    res[path] = (out, dt * nq / m)
    print("path", path, "last_path", eng.profile_read()["last_path"], "ms per 1024-query batch", round(dt * nq / m * 1e3, 1))
a, b = res[2][0], res[1][0]
print("first 64 queries bit-identical:", bool(np.array_equal(a[0][:64], b[0]) and np.array_equal(a[1][:64].view(np.uint32), b[1].view(np.uint32))))
PY
