"""The N>1 path's host logic on CPU: world_size-2 gloo run (no GPU).

What is exercised: the product's shard partition (mvf_shard_range), the packed (key << 32 | LOCAL row
id) candidate words and their order-preserving key map (mvf_ordered_key_f32), one all_gather of nq*k
words per rank, and the cross-rank merge rule of kernels.cuh (rank_word / emit_result): equal keys
order by (rank, position in the rank's sorted list) and the global row is list_base[rank] + local id
in 64 bits -- so the space may sit beyond row 2^32.  The per-shard top-k itself comes from the CPU
oracle here (the CUDA kernels are covered by -m gpu tests); the claim under test is that the sharded
pipeline returns exactly the unsharded result.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import metrovector_b200 as mv
import oracle


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _pack(keys_u32, ids_u64):
    return ((keys_u32.astype(np.uint64) << np.uint64(32)) | ids_u64.astype(np.uint64))


INVALID = np.uint64(0xFFFFFFFFFFFFFFFF)


def _worker(rank, world, port, dtype, metric, n, d, k, nq, out_dir, offset=0):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        begin, count = mv.shard_range(n, world, rank)
        shard = oracle.synth(dtype, 42, begin, count, d)            # this rank regenerates only its rows
        queries = oracle.synth(dtype, 43, 0, nq, d)
        local = np.full((nq, k), np.uint64(0xFFFFFFFFFFFFFFFF), dtype=np.uint64)
        for qi in range(nq):
            q = queries[qi].astype(np.float32) if dtype in (oracle.F32, oracle.F16) else queries[qi]
            if count:
                ids, sc, isc = oracle.topk(shard, dtype, metric, q, k, row_base=begin + offset)
                ids = ids - np.uint64(begin + offset)               # what travels is the row's position inside the shard
                if dtype in (oracle.I8, oracle.U8) and metric != oracle.COS:
                    v = isc.astype(np.int64) if metric == oracle.L2 else -isc.astype(np.int64)
                    keys = (v + 2 ** 31).astype(np.uint32)
                else:
                    rank_val = -sc if metric == oracle.IP else sc
                    keys = np.array([mv.ordered_key_f32(x) for x in rank_val], dtype=np.uint32)
                local[qi, :len(ids)] = _pack(keys, ids)
        t = torch.from_numpy(local.view(np.int64).copy())
        gathered = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)                                 # the one exchange step
        allc = np.stack([g.numpy().view(np.uint64) for g in gathered])   # [world, nq, k]
        bases = [mv.shard_range(n, world, r)[0] + offset for r in range(world)]
        # ranked words: the low half names (rank, position) instead of the local id, which it is resolved back to
        pos = (np.arange(world, dtype=np.uint64)[:, None, None] << np.uint64(16)) | np.arange(k, dtype=np.uint64)[None, None, :]
        ranked = np.where(allc == INVALID, INVALID, (allc & np.uint64(0xFFFFFFFF00000000)) | pos)
        top = np.sort(ranked.transpose(1, 0, 2).reshape(nq, world * k), axis=1)[:, :k]
        keys = np.full((nq, k), np.uint32(0xFFFFFFFF), dtype=np.uint32)
        gids = np.full((nq, k), INVALID, dtype=np.uint64)
        for qi in range(nq):
            for j in range(k):
                w = int(top[qi, j])
                if w == int(INVALID):
                    continue
                r, p = (w >> 16) & 0xFFFF, w & 0xFFFF
                keys[qi, j] = w >> 32
                gids[qi, j] = np.uint64(bases[r] + (int(allc[r, qi, p]) & 0xFFFFFFFF))
        np.save(os.path.join(out_dir, "keys_%d.npy" % rank), keys)
        np.save(os.path.join(out_dir, "gids_%d.npy" % rank), gids)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dtype,metric,n,d,k,offset", [
    (oracle.F32, oracle.L2, 1001, 24, 10, 0),
    (oracle.U8, oracle.L2, 777, 16, 25, 0),          # tie-rich integers: lowest global index must win across shards
    (oracle.F16, oracle.COS, 640, 32, 10, 0),
    (oracle.I8, oracle.IP, 3, 8, 5, 0),              # fewer rows than k, one rank nearly empty
    (oracle.U8, oracle.L2, 777, 16, 25, (1 << 33) + 5),   # the same ties in a space placed beyond row 2^32
])
def test_two_rank_sharded_search_equals_unsharded(tmp_path, dtype, metric, n, d, k, offset):
    world, nq = 2, 3
    port = _free_port()
    mp.spawn(_worker, args=(world, port, dtype, metric, n, d, k, nq, str(tmp_path), offset), nprocs=world, join=True)
    db = oracle.synth(dtype, 42, 0, n, d)
    queries = oracle.synth(dtype, 43, 0, nq, d)
    k0, k1 = np.load(tmp_path / "keys_0.npy"), np.load(tmp_path / "keys_1.npy")
    g0, g1 = np.load(tmp_path / "gids_0.npy"), np.load(tmp_path / "gids_1.npy")
    assert np.array_equal(k0, k1) and np.array_equal(g0, g1)         # every rank holds the same answer
    for qi in range(nq):
        q = queries[qi].astype(np.float32) if dtype in (oracle.F32, oracle.F16) else queries[qi]
        ids, sc, isc = oracle.topk(db, dtype, metric, q, k, row_base=offset)
        valid = g0[qi] != INVALID
        assert int(valid.sum()) == len(ids)
        assert g0[qi][valid].tolist() == ids.tolist()
        if dtype in (oracle.F32, oracle.F16) or metric == oracle.COS:
            back = np.array([mv.ordered_key_to_f32(int(x)) for x in k0[qi][valid]], dtype=np.float32)
            want = -sc if metric == oracle.IP else sc
            assert np.array_equal(back.view(np.uint32), (want + np.float32(0)).view(np.uint32))


def test_shard_ranges_partition_the_rows():
    for n in (0, 1, 7, 8, 9, 1000003):
        for world in (1, 2, 3, 4, 8):
            covered = 0
            for r in range(world):
                b, c = mv.shard_range(n, world, r)
                assert b == min(covered, n) or c == 0
                covered += c
            assert covered == n
    with pytest.raises(mv.Build):
        mv.shard_range(10, 2, 2)


def test_every_rank_takes_the_same_scan_or_tensor_decision():
    # the choice used to be made from a rank's own row count; with n = world * 4096 - 1 rank 0 (4096
    # rows) took the tensor path and the tail rank (4095) the scan, and the ranks exchanged keys in two encodings.  The rule
    # now: decide on the smallest non-empty shard of mvf_shard_range -- a function of (global_n, world) only.
    import metrovector_b200 as mv
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, world - 1, world, world * 4096 - 1, world * 4096, world * 4096 + 1, 100_000_000, 999_999_937):
            if n < 0:
                continue
            rows, _ = mv.path_decision(n, world, 1, 2, 64, 10, 768)
            shards = [mv.shard_range(n, world, r)[1] for r in range(world)]
            nonempty = [c for c in shards if c]
            assert rows == (min(nonempty) if nonempty else 0), (world, n, rows, shards)
    # Float16, 8 queries: 4096 rows per rank qualify, a 4095-row tail shard takes every rank to the scan kernel
    assert mv.path_decision(2 * 4096, 2, 1, 2, 8, 10, 64) == (4096, True)
    assert mv.path_decision(2 * 4096 - 1, 2, 1, 2, 8, 10, 64) == (4095, False)
    assert mv.path_decision(8 * 4096 - 1, 8, 3, 0, 16, 10, 128) == (4095, False)
    # what the tensor path does not take in auto mode: small batches, k > 512, Float32 rows beyond 1024 components
    assert mv.path_decision(10**8, 8, 1, 2, 3, 10, 768)[1] is False
    assert mv.path_decision(10**8, 8, 1, 2, 1024, 129, 768)[1] is True      # k up to 512 stays on the tensor path
    assert mv.path_decision(10**8, 8, 1, 2, 1024, 513, 768)[1] is False
    assert mv.path_decision(10**8, 8, 3, 2, 1024, 10, 128)[1] is True       # integer cosine: exact keys for the candidates
    assert mv.path_decision(10**6, 1, 0, 0, 1024, 10, 768)[1] is True
    assert mv.path_decision(10**6, 1, 0, 0, 1024, 10, 2048)[1] is False
