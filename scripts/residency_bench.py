#!/usr/bin/env python
"""Residency at config-B size: write a 1M x 768 f32 MVF file, open it through the reader, upload the
mapped block to HBM (both paths), report GB/s, and check that searching the uploaded space gives
exactly what the device-generated space gives (same generator -> same rows -> same top-k)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import metrovector_b200 as mv  # noqa: E402

n, d = int(os.environ.get("ROWS", 1_000_000)), 768
path = "/tmp/residency_b.mvf"
t0 = time.time()
rows = mv.synth_fill_host(0, 0x4D564631, 0, n, d)
t1 = time.time()
b = mv.MvfBuilder()
b.add_vector_space("embeddings", d, mv.VectorType.Dense, mv.DistanceMetric.L2, mv.DataType.Float32)
b.add_vectors_raw("embeddings", rows)
b.save(path)
t2 = time.time()
eng = mv.GpuEngine(0)
q = mv.synth_fill_host(0, 0x4D564632, 0, 4, d)
ref_space = eng.synthetic(mv.DataType.Float32, mv.DistanceMetric.L2, n, d, 0x4D564631)
ref = ref_space.search(q, 10)
out = {"rows": n, "file_gb": os.path.getsize(path) / 1e9, "host_fill_s": t1 - t0, "write_s": t2 - t1, "uploads": []}
# every residency variant in turn (MVF_HOST_REGISTER): what the driver accepts, and how fast each is
for variant in ("auto", "ro", "rw", "private", "none", "none"):
    if variant == "auto":
        os.environ.pop("MVF_HOST_REGISTER", None)
    else:
        os.environ["MVF_HOST_REGISTER"] = variant
    r = mv.MvfReader.open(path)
    r.validate()
    t3 = time.time()
    gs = eng.upload(r.vector_space("embeddings"))
    t4 = time.time()
    st = gs.upload_stats()
    crc, crc_ms = gs.checksum()
    gs.verify_checksum(r.vector_space("embeddings"))
    got = gs.search(q, 10)
    same = bool(np.array_equal(got[0], ref[0]) and np.array_equal(got[1].view(np.uint32), ref[1].view(np.uint32)))
    rows_ok = bool(np.array_equal(gs.read_rows(n - 3, 3), rows[-3:]))
    out["uploads"].append({"variant": variant, "path": {1: "cudaHostRegister'd mapping, DMA in place", 2: "pinned staging"}[st["path"]],
                           "seconds": st["seconds"], "gb_per_s": st["bytes"] / 1e9 / st["seconds"], "wall_s": t4 - t3,
                           "search_identical_to_generated_space": same, "tail_rows_identical": rows_ok, "note": st["note"], "device_crc32_ms": crc_ms, "device_crc32_gb_per_s": st["bytes"] / 1e9 / (crc_ms * 1e-3)})
    gs.free()
    r.close()
os.remove(path)
print(json.dumps(out))
