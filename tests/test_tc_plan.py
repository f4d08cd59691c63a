"""Planner of the tensor-core kernel (tc_path.hpp::make_plan) checked on the CPU: whatever shape the host picks has
to fit the hardware limits the kernel is written against, for every problem size the path accepts."""
import itertools

import pytest

import metrovector_b200 as mv

SMEM_LIMIT = 227 * 1024


def _check(pl, nq, k, n, d, es, sm):
    cols = 256 // pl["epi_groups"]
    assert pl["mtiles"] == -(-nq // 128) and pl["ntiles"] == -(-n // 256)
    assert pl["kblocks"] == -(-(d * es) // 128)
    # shared memory, ring depth, epilogue shape
    assert pl["smem"] <= SMEM_LIMIT, pl
    assert pl["epi_groups"] in (2, 4) and 2 <= pl["stages"] <= 6
    if es == 4:                                                              # split tf32 (K2'): two planes per operand
        assert pl["epi_groups"] == 2 and pl["pair"] == (pl["cluster"] == 2)
        assert pl["stages"] >= (3 if pl["pair"] and pl["cap"] < 512 else 2), pl
    else:
        assert (pl["stages"] >= 3) or pl["cap"] >= 512, pl                   # a 2-deep ring only under the largest sort scratches
    # candidate buffers: power of two, room for one more tile past the trigger, a compaction leaves k
    assert pl["cap"] in (128, 256, 512, 1024)
    assert k <= pl["trigger"] <= pl["cap"] - cols, pl
    # work partition: runs cover the tiles exactly, every cluster has work, grid fits the chip
    assert pl["chunks"] >= 1 and pl["chunks"] * pl["tiles_per_chunk"] >= pl["ntiles"]
    assert (pl["chunks"] - 1) * pl["tiles_per_chunk"] < pl["ntiles"]
    assert pl["cluster"] in (1, 2) and (pl["cluster"] == 2) == (pl["mtiles"] >= 2)
    assert pl["mgroups"] == -(-pl["mtiles"] // pl["cluster"]) and pl["items"] == pl["mgroups"] * pl["chunks"]
    assert pl["grid"] % pl["cluster"] == 0 and pl["cluster"] <= pl["grid"] <= sm
    assert pl["grid"] // pl["cluster"] <= pl["items"]
    assert not pl["pair"] or pl["cluster"] == 2
    # pacing only where several query-tile groups share rows
    assert (pl["sync_window"] == 0) == (pl["mgroups"] < 2)
    assert pl["sync_window"] <= 64


@pytest.mark.parametrize("dtype,es", [(0, 4), (1, 2), (2, 1), (3, 1)])
def test_plan_invariants_over_the_accepted_problem_space(dtype, es):
    nqs = (1, 3, 64, 128, 129, 256, 300, 1024, 4096, 100000)
    ks = (1, 10, 16, 17, 32, 33, 56, 57, 64, 100, 112, 113, 120, 128, 129, 200, 256, 257, 384, 500, 512)
    ns = (1, 255, 256, 257, 4096, 100_000, 10_000_000, 100_000_000, (1 << 32) - 1)
    ds = (1, 16, 64, 96, 128, 129, 256, 300, 768, 2048, 4096)
    for nq, k, n, d in itertools.product(nqs, ks, ns, ds):
        if dtype == 0 and d > 1024:
            # split tf32: the tensor core's f32 accumulation over 3 d / 8 MMA steps eats the stated tolerance beyond ~1k
            # components (DESIGN.md section 4, K2'): longer Float32 rows stay on the scan kernel
            with pytest.raises(mv.Build):
                mv.tc_plan(dtype, nq, k, n, d)
            continue
        if dtype == 0 and nq <= 128 and k > 56:
            # one query tile of a Float32 space runs single CTAs with 96 KiB stages: no room for the 512-entry sort
            # scratch, the planner refuses and the engine keeps such searches on the scan kernel
            with pytest.raises(mv.Build):
                mv.tc_plan(dtype, nq, k, n, d)
            continue
        _check(mv.tc_plan(dtype, nq, k, n, d), nq, k, n, d, es, 148)
    for sm in (2, 16, 132, 148, 160):                                         # other chip sizes
        for nq, k, n, d in itertools.product((130, 4096), (10, 128), (1000, 10_000_000), (128, 768)):
            _check(mv.tc_plan(dtype, nq, k, n, d, sm), nq, k, n, d, es, sm)


def test_plan_of_the_baseline_configs():
    c = mv.tc_plan(1, 4096, 100, 10_000_000, 768)                            # config C
    assert (c["mtiles"], c["mgroups"], c["chunks"], c["items"], c["grid"]) == (32, 16, 37, 592, 148)
    assert c["pair"] == 1 and c["epi_groups"] == 2 and c["cap"] == 512 and c["trigger"] == 240 and c["sync_window"] == 9
    d = mv.tc_plan(1, 1024, 10, 100_000_000, 768)                            # config D, batch 1024
    assert (d["mgroups"], d["chunks"], d["items"]) == (4, 148, 592) and d["trigger"] == 48 and d["cap"] == 256
    assert d["pair"] == 1 and d["stages"] == 6 and d["sync_window"] == 2
    e = mv.tc_plan(3, 1024, 10, 1_000_000_000, 128)                          # config E, batch 1024
    assert e["pair"] == 0 and e["epi_groups"] == 4 and e["cap"] == 128 and e["trigger"] == 48 and e["kblocks"] == 1


def test_plan_of_a_float32_batch():
    b = mv.tc_plan(0, 1024, 10, 1_000_000, 768)                              # config B's space, batch 1024 (K2')
    assert b["pair"] == 1 and b["cluster"] == 2 and b["epi_groups"] == 2 and b["kblocks"] == 24 and b["stages"] == 3
    assert b["smem"] <= SMEM_LIMIT


def test_plan_rejects_what_the_path_does_not_take():
    for args in ((4, 10, 10, 1000, 128), (1, 10, 513, 1000, 128), (1, 0, 10, 1000, 128), (1, 10, 10, 0, 128), (0, 10, 100, 1000, 128)):
        with pytest.raises(mv.Build):
            mv.tc_plan(*args)
