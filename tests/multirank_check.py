"""Row-sharded search on N GPUs == the unsharded search == the CPU oracle.

Run under torchrun (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multirank_check.py
Each rank holds its row shard (device-generated), every rank calls the same search, results are
exchanged inside the library: nq*k packed candidates per rank, over peer memory (small payloads) or ONE ncclAllGather.  Rank 0 additionally
holds the whole database on a second, communicator-less engine and checks bit-identity.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import metrovector_b200 as mv  # noqa: E402
import oracle  # noqa: E402
from util import as_query, check_batch  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = mv.GpuEngine(local)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(mv.GpuEngine.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    eng.comm_init(uid.cpu().numpy().tobytes(), rank, world)
    solo = mv.GpuEngine(local) if rank == 0 else None
    failures = 0
    cases = [(0, 0, 20000, 128, 10, 3), (1, 2, 30001, 96, 10, 2), (3, 0, 50000, 128, 25, 4), (2, 1, 7, 16, 5, 2),
             (3, 2, 40000, 64, 10, 1), (0, 1, 1000, 768, 100, 9),
             # large k inside the peer-exchange payload (<= 512 candidates): fused into the scan launch / its own kernel
             (3, 0, 30000, 64, 500, 1), (0, 2, 20000, 32, 256, 2),
             # batches that take the tensor-core path on every rank (keys are K2 ranking values until the final merge)
             (3, 0, 60000, 128, 10, 40), (1, 2, 50000, 128, 10, 70), (1, 0, 50000, 64, 33, 130), (2, 1, 40000, 96, 10, 16),
             # Float32 batches: the split-tf32 form (K2')
             (0, 0, 60000, 96, 10, 48), (0, 2, 50000, 128, 10, 200),
             # integer cosine on the tensor path: its candidates carry final keys (nothing to decode after the exchange)
             (3, 2, 50000, 128, 10, 40), (2, 2, 40000, 64, 33, 130),
             # the tail shard falls under the tensor path's minimum (4096 rows) while the others do not: every rank must
             # make the same choice or the exchanged keys are in two encodings
             (1, 2, world * 4096 - 1, 64, 10, 8), (3, 0, world * 4096 - 1, 128, 10, 16)]
    for dtype, metric, n, d, k, nq in cases:
        gs = eng.synthetic(mv.DataType(dtype), mv.DistanceMetric(metric), n, d, seed=321)
        b, c = mv.shard_range(n, world, rank)
        assert (gs.row_base, gs.local_n, gs.global_n) == (b, c, n)
        q = mv.synth_fill_host(dtype, 654, 0, nq, d)
        queries = q if (dtype == 1 and nq >= 4) else np.stack([as_query(dtype, r) for r in q])   # f16 batches stay f16 (tensor path)
        ids, sc, isc = gs.search(queries, k)
        # every rank must hold the same answer
        t = torch.from_numpy(ids.view(np.int64).copy()).cuda()
        ref = t.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(t, ref), "ranks disagree"
        if rank == 0:
            db = mv.synth_fill_host(dtype, 321, 0, n, d)
            check_batch(db, dtype, metric, queries.astype(np.float32) if dtype == 1 else queries, k, ids, sc, isc)
            full = solo.from_host(db, mv.DataType(dtype), mv.DistanceMetric(metric))
            # tail-shard cases: the sharded ranks stay on the scan kernel (4095-row tail); compare like with like
            solo.set_path(1 if n == world * 4096 - 1 else 0)
            ids1, sc1, isc1 = full.search(queries, k)
            solo.set_path(0)
            same = np.array_equal(ids, ids1) and np.array_equal(sc.view(np.uint32), sc1.view(np.uint32)) and np.array_equal(isc, isc1)
            if not same:
                failures += 1
                print("MISMATCH sharded vs single", (dtype, metric, n, d, k, nq))
            full.free()
        gs.free()
    # forced tensor path with empty tail shards (fewer rows than ranks)
    eng.set_path(2)
    if solo is not None:
        solo.set_path(2)
    for dtype, n, d, nq in ((3, 1, 128, 9), (1, max(1, world - 1), 64, 5)):
        gs = eng.synthetic(mv.DataType(dtype), mv.DistanceMetric.L2, n, d, seed=11)
        q = mv.synth_fill_host(dtype, 12, 0, nq, d)
        ids, sc, isc = gs.search(q, 4)
        if rank == 0:
            db = mv.synth_fill_host(dtype, 11, 0, n, d)
            full = solo.from_host(db, mv.DataType(dtype), mv.DistanceMetric.L2)
            ids1, sc1, isc1 = full.search(q, 4)
            if not (np.array_equal(ids, ids1) and np.array_equal(sc.view(np.uint32), sc1.view(np.uint32))):
                failures += 1
                print("MISMATCH forced tensor path with empty shards", (dtype, n))
            full.free()
        gs.free()
    eng.set_path(0)
    if solo is not None:
        solo.set_path(0)
    # block checksum of a sharded upload: every rank hashes its rows, the shard CRCs are combined (collective call)
    import tempfile
    import zlib
    n, d = 20000, 40
    dbf = mv.synth_fill_host(1, 91, 0, n, d)
    path = os.path.join(tempfile.gettempdir(), "mvf_multirank_%d.mvf" % os.getppid())
    if local == 0:
        b = mv.MvfBuilder()
        b.add_vector_space("s", d, mv.VectorType.Dense, mv.DistanceMetric.Cosine, mv.DataType.Float16)
        b.add_vectors_raw("s", dbf)
        b.save(path)
    dist.barrier()
    r = mv.MvfReader.open(path)
    sp = r.vector_space("s")
    gs = eng.upload(sp)
    crc, _ = gs.checksum()
    gs.verify_checksum(sp)
    if crc != (zlib.crc32(dbf.tobytes()) & 0xffffffff):
        failures += 1
        print("MISMATCH sharded checksum on rank", rank)
    gs.free()
    r.close()
    dist.barrier()
    if local == 0:
        os.remove(path)
    # external ids and a global row mask on a sharded space: ids table replicated, mask re-based per shard
    n, d, k = 30000, 64, 10
    gs = eng.synthetic(mv.DataType.UInt8, mv.DistanceMetric.L2, n, d, seed=77)
    ext = (np.arange(n, dtype=np.uint64)[::-1] * 7 + 5).copy()
    excluded = (np.arange(n) % 3) == 0
    gs.set_ids(ext)
    gs.set_row_mask(excluded)
    q = mv.synth_fill_host(3, 78, 0, 20, d)
    ids, sc, isc = gs.search(q, k)
    if rank == 0:
        db = mv.synth_fill_host(3, 77, 0, n, d)
        keep = np.flatnonzero(~excluded)
        for qi in range(20):
            o_ids, _, o_isc = oracle.topk(db[keep], 3, 0, q[qi], k)
            want = ext[keep[o_ids.astype(np.int64)]]
            if ids[qi].tolist() != want.tolist() or isc[qi].tolist() != o_isc.tolist():
                failures += 1
                print("MISMATCH external ids / mask on shards, query", qi)
    gs.free()
    # shards of a space of more than 2^32 rows (caller-cut shards, row_base far beyond 32 bits): local ids travel through the
    # exchange, the cross-rank merge orders ties by (rank, position) and adds every rank's first row in 64 bits
    big = (1 << 34) + 99
    for dtype, metric, n, d, k, nq in ((3, 0, 40000, 64, 10, 2), (1, 2, 50000, 64, 10, 70), (3, 0, 9000, 16, 100, 9), (2, 1, 30000, 32, 10, 1)):
        db = mv.synth_fill_host(dtype, 31, 0, n, d)
        if dtype == 3 and d == 16:
            db[:] = db % 3                                     # thousands of exact ties across the shard boundary
        b, c = mv.shard_range(n, world, rank)
        gs = eng.from_host(db[b:b + c], mv.DataType(dtype), mv.DistanceMetric(metric), row_base=big + b, global_n=big + n)
        q = mv.synth_fill_host(dtype, 32, 0, nq, d)
        if dtype == 3 and d == 16:
            q[:] = q % 3
        queries = q if (dtype == 1 and nq >= 4) else np.stack([as_query(dtype, r) for r in q])
        ids, sc, isc = gs.search(queries, k)
        if rank == 0:
            full = solo.from_host(db, mv.DataType(dtype), mv.DistanceMetric(metric), row_base=big, global_n=big + n)
            ids1, sc1, isc1 = full.search(queries, k)
            if not (np.array_equal(ids, ids1) and np.array_equal(sc.view(np.uint32), sc1.view(np.uint32)) and np.array_equal(isc, isc1)
                    and ids.min() >= big):
                failures += 1
                print("MISMATCH beyond 2^32 rows", (dtype, metric, n, d, k, nq))
            full.free()
        gs.free()
    # back-to-back searches without a host sync in between (device-resident queries and results): the peer-memory
    # exchange alternates two buffers and relies on nobody running two searches ahead of a reader -- 400 searches of
    # 1..3 queries queued at once, every result compared with the communicator-less engine afterwards
    n, d, k, rounds = 24000, 64, 10, 400
    gs = eng.synthetic(mv.DataType.Float16, mv.DistanceMetric.Cosine, n, d, seed=5)
    qh = mv.synth_fill_host(1, 6, 0, 3 * rounds, d).astype(np.float32)
    dq = torch.from_numpy(qh).cuda()
    d_ids = torch.zeros((rounds, 3, k), dtype=torch.int64, device="cuda")
    d_sc = torch.zeros((rounds, 3, k), dtype=torch.float32, device="cuda")
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    for i in range(rounds):
        nq = 1 + i % 3
        gs.search_device(dq[3 * i].data_ptr(), 0, nq, k, 2, d_ids[i].data_ptr(), d_sc[i].data_ptr())
    eng.synchronize()
    torch.cuda.synchronize()
    if rank == 0:
        full = solo.from_host(mv.synth_fill_host(1, 5, 0, n, d), mv.DataType.Float16, mv.DistanceMetric.Cosine)
        ids1, sc1, _ = full.search(qh, k)
        got_i, got_s = d_ids.cpu().numpy().view(np.uint64), d_sc.cpu().numpy()
        for i in range(rounds):
            nq = 1 + i % 3
            if not (np.array_equal(got_i[i, :nq], ids1[3 * i:3 * i + nq]) and
                    np.array_equal(got_s[i, :nq].view(np.uint32), sc1[3 * i:3 * i + nq].view(np.uint32))):
                failures += 1
                print("MISMATCH back-to-back search", i)
                break
        full.free()
    gs.free()
    info = eng.exchange_info()
    if rank == 0:
        print("exchange:", info)
    if os.environ.get("MVF_PEER_EXCHANGE", "1") != "0":
        # both routes must have been exercised: small payloads over peer memory, large batches over NCCL
        if not (info["peer_ready"] and info["peer_searches"] > rounds and info["nccl_searches"] > 0):
            failures += 1
            print("peer exchange not in use on rank", rank, info)
    dist.barrier()
    if rank == 0:
        print("multirank ok" if failures == 0 else "multirank FAILED", "world", world)
    eng.shutdown()
    dist.destroy_process_group()
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
