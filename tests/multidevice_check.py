"""ONE process, every visible GPU: mvf_gpu_init_multi == the single-GPU answer == the CPU oracle.

    python tests/multidevice_check.py          # needs >= 2 GPUs; prints "multidevice ok world N"

The group engine row-shards every space over its devices, runs the shards' kernels concurrently, exchanges the local
top-k (peer stores for small payloads, else one ncclAllGather per device) and merges on device 0.  A second, single-device
engine on GPU 0 holds the whole database for the bit-identity check.
"""
import os
import sys
import tempfile
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import metrovector_b200 as mv  # noqa: E402
import oracle  # noqa: E402
from util import as_query, check_batch  # noqa: E402


def main():
    ndev = torch.cuda.device_count()
    if ndev < 2:
        print("multidevice check needs >= 2 GPUs, found", ndev)
        return 2
    eng = mv.GpuEngine(list(range(ndev)))
    assert eng.device_count() == ndev
    solo = mv.GpuEngine(0)
    failures = 0

    def same(a, b):
        return np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32)) and np.array_equal(a[2], b[2])

    cases = [(0, 0, 20000, 128, 10, 3), (1, 2, 30001, 96, 10, 2), (3, 0, 50000, 128, 25, 4), (2, 1, 7, 16, 5, 2),
             (3, 2, 40000, 64, 10, 1), (0, 1, 1000, 768, 100, 9),
             # large k inside the peer-exchange payload (<= 512 candidates): fused into the scan launch / its own kernel
             (3, 0, 30000, 64, 500, 1), (0, 2, 20000, 32, 256, 2),
             # batches that take the tensor-core path on every shard
             (3, 0, 60000, 128, 10, 40), (1, 2, 50000, 128, 10, 70), (1, 0, 50000, 64, 33, 130), (2, 1, 40000, 96, 10, 16),
             (0, 0, 60000, 96, 10, 48), (0, 2, 50000, 128, 10, 200),
             # integer cosine on the tensor path: its candidates carry final keys (nothing to decode after the exchange)
             (3, 2, 50000, 128, 10, 40), (2, 2, 40000, 64, 33, 130),
             # the tail shard falls under the tensor path's minimum while the others do not: one choice for all shards
             (1, 2, ndev * 4096 - 1, 64, 10, 8), (3, 0, ndev * 4096 - 1, 128, 10, 16)]
    for dtype, metric, n, d, k, nq in cases:
        gs = eng.synthetic(mv.DataType(dtype), mv.DistanceMetric(metric), n, d, seed=321)
        assert (gs.local_n, gs.global_n) == (n, n)
        q = mv.synth_fill_host(dtype, 654, 0, nq, d)
        queries = q if (dtype == 1 and nq >= 4) else np.stack([as_query(dtype, r) for r in q])
        got = gs.search(queries, k)
        db = mv.synth_fill_host(dtype, 321, 0, n, d)
        check_batch(db, dtype, metric, queries.astype(np.float32) if dtype == 1 else queries, k, *got)
        full = solo.from_host(db, mv.DataType(dtype), mv.DistanceMetric(metric))
        # tail-shard cases: the shards stay on the scan kernel (4095-row tail); compare like with like
        solo.set_path(1 if n == ndev * 4096 - 1 else 0)
        want = full.search(queries, k)
        solo.set_path(0)
        if not same(got, want):
            failures += 1
            print("MISMATCH group vs single", (dtype, metric, n, d, k, nq))
        full.free()
        gs.free()

    # forced tensor path with empty tail shards (fewer rows than devices)
    eng.set_path(2)
    solo.set_path(2)
    for dtype, n, d, nq in ((3, 1, 128, 9), (1, ndev - 1, 64, 5)):
        db = mv.synth_fill_host(dtype, 11, 0, n, d)
        gs = eng.from_host(db, mv.DataType(dtype), mv.DistanceMetric.L2)
        full = solo.from_host(db, mv.DataType(dtype), mv.DistanceMetric.L2)
        q = mv.synth_fill_host(dtype, 12, 0, nq, d)
        got, want = gs.search(q, 4), full.search(q, 4)
        if not same(got, want):
            failures += 1
            print("MISMATCH forced tensor path, empty shards", (dtype, n))
        gs.free()
        full.free()
    eng.set_path(0)
    solo.set_path(0)

    # a real file: tombstones, external ids, block checksum (per-shard CRCs combined), payload gather
    n, d, k = 30000, 48, 10
    db = mv.synth_fill_host(1, 77, 0, n, d)
    deleted = (np.arange(n) % 5) == 0
    ext = (np.arange(n, dtype=np.uint64)[::-1] * 3 + 1).copy()
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "g.mvf")
        b = mv.MvfBuilder()
        b.add_vector_space("s", d, mv.VectorType.Dense, mv.DistanceMetric.Cosine, mv.DataType.Float16)
        b.add_vectors_raw("s", db)
        b.set_tombstones("s", deleted)
        b.set_vector_ids("s", ext)
        b.save(path)
        r = mv.MvfReader.open(path)
        sp = r.vector_space("s")
        gs = eng.upload(sp)
        crc, _ = gs.checksum()
        if crc != (zlib.crc32(db.tobytes()) & 0xffffffff):
            failures += 1
            print("MISMATCH sharded checksum")
        gs.verify_checksum(sp)
        q = mv.synth_fill_host(1, 78, 0, 6, d).astype(np.float32)
        ids, sc, isc = gs.search(q, k)
        keep = np.flatnonzero(~deleted)
        for qi in range(6):
            o_ids, o_sc, _ = oracle.topk(db[keep], 1, 2, q[qi], k)
            want = ext[keep[o_ids.astype(np.int64)]]
            if ids[qi].tolist() != want.tolist():
                failures += 1
                print("MISMATCH tombstones / external ids on the group engine, query", qi)
        pos = np.array([0, n // ndev, n - 1, 12345 % n], dtype=np.uint64)
        if not np.array_equal(gs.get_vectors(pos), db[pos.astype(np.int64)]):
            failures += 1
            print("MISMATCH payload gather across shards")
        if not np.array_equal(gs.read_rows(n // ndev - 3, 7), db[n // ndev - 3:n // ndev + 4]):
            failures += 1
            print("MISMATCH read_rows across a shard boundary")
        gs.free()
        r.close()

    # a space placed beyond 2^32 rows: local ids travel, member 0's merge adds every shard's first row in 64 bits
    big = (1 << 34) + 99
    for dtype, metric, n, d, k, nq in ((3, 0, 40000, 64, 10, 2), (1, 2, 50000, 64, 10, 70), (3, 0, 9000, 16, 100, 9)):
        db = mv.synth_fill_host(dtype, 31, 0, n, d)
        q = mv.synth_fill_host(dtype, 32, 0, nq, d)
        if d == 16:
            db[:] = db % 3                                         # exact ties across shard boundaries
            q[:] = q % 3
        queries = q if (dtype == 1 and nq >= 4) else np.stack([as_query(dtype, r) for r in q])
        gs = eng.from_host(db, mv.DataType(dtype), mv.DistanceMetric(metric), row_base=big, global_n=big + n)
        full = solo.from_host(db, mv.DataType(dtype), mv.DistanceMetric(metric), row_base=big, global_n=big + n)
        got, want = gs.search(queries, k), full.search(queries, k)
        if not (same(got, want) and got[0].min() >= big):
            failures += 1
            print("MISMATCH beyond 2^32 rows", (dtype, metric, n, d, k, nq))
        gs.free()
        full.free()

    # back-to-back searches with device-resident queries/results on device 0 (no host sync in between): the peer-memory
    # exchange alternates two buffers; 300 queued searches, each compared with the single-device engine afterwards
    n, d, k, rounds = 24000, 64, 10, 300
    db = mv.synth_fill_host(1, 5, 0, n, d)
    gs = eng.from_host(db, mv.DataType.Float16, mv.DistanceMetric.Cosine)
    full = solo.from_host(db, mv.DataType.Float16, mv.DistanceMetric.Cosine)
    qh = mv.synth_fill_host(1, 6, 0, 3 * rounds, d).astype(np.float32)
    with torch.cuda.device(0):
        dq = torch.from_numpy(qh).cuda()
        d_ids = torch.zeros((rounds, 3, k), dtype=torch.int64, device="cuda")
        d_sc = torch.zeros((rounds, 3, k), dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        for i in range(rounds):
            gs.search_device(dq[3 * i].data_ptr(), 0, 1 + i % 3, k, 2, d_ids[i].data_ptr(), d_sc[i].data_ptr())
        eng.synchronize()
        got_i, got_s = d_ids.cpu().numpy().view(np.uint64), d_sc.cpu().numpy()
    ids1, sc1, _ = full.search(qh, k)
    for i in range(rounds):
        nq = 1 + i % 3
        if not (np.array_equal(got_i[i, :nq], ids1[3 * i:3 * i + nq]) and
                np.array_equal(got_s[i, :nq].view(np.uint32), sc1[3 * i:3 * i + nq].view(np.uint32))):
            failures += 1
            print("MISMATCH back-to-back search", i)
            break
    gs.free()
    full.free()
    info = eng.exchange_info()
    print("exchange:", info)
    if os.environ.get("MVF_PEER_EXCHANGE", "1") != "0" and not (info["peer_ready"] and info["peer_searches"] > rounds and info["nccl_searches"] > 0):
        failures += 1
        print("peer exchange not in use", info)

    # single-device-only calls answer Build on a group engine
    try:
        eng.set_stream(0)
        failures += 1
        print("set_stream on a group engine did not fail")
    except mv.Build:
        pass
    eng.shutdown()
    solo.shutdown()
    print("multidevice ok" if failures == 0 else "multidevice FAILED", "world", ndev)
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
