#!/usr/bin/env python
"""Regenerates the committed fixtures in tests/golden/.

Run from the repo root:  python tests/golden/make_golden.py
Inputs are the datasets hard-coded in the reference's example programs and unit-test fixture
(/root/reference/examples/simple.rs:15-21, examples/similarity_search.rs:43-76,
src/tests/test_utils.rs:52-76) restated here, written through the pure-Python MVF writer in
oracle/mvf_format.py, and searched with the CPU oracle (oracle/oracle.c).  The reference itself
is Rust and cannot run in this image, so these are oracle-produced vectors cross-checked against
the hand-derived values of SURVEY.md Appendix B (tests/test_oracle_golden.py holds those).
"""
import json
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle import mvf_format as fmt  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
f = np.float32


def simple_rows():
    return np.array([[1, 2, 3, 4], [5, 6, 7, 8], [9, 10, 11, 12], [2, 4, 6, 8], [1, 3, 5, 7]], dtype=f)


def clustered_rows():
    rows = []
    for base in ((1, 1, 1, 1), (5, 5, 5, 5)):
        for i in range(20):
            n = f(i) * f(0.1)
            rows.append([f(base[0]) + n, f(base[1]) - n, f(base[2]) + n * f(0.5), f(base[3]) - n * f(0.5)])
    for i in range(20):
        n = f(i) * f(0.1)
        rows.append([f(-2.0) + n, f(3.0) - n, f(0.0) + n, f(4.0) - n * f(0.5)])
    return np.array(rows, dtype=f)


def bits(x):
    return [int(struct.unpack("<I", struct.pack("<f", float(v)))[0]) for v in x]


def write(name, data):
    with open(os.path.join(OUT, name), "wb") as fh:
        fh.write(data)


def main():
    known = {}
    # 1. examples/simple.rs
    rows = simple_rows()
    write("simple_f32.mvf", fmt.build_mvf([dict(name="embeddings", dimension=4, vector_type=0,
                                                 distance_metric=0, data_type=0, data=rows.tobytes())]))
    q = np.array([2.5, 4.5, 6.5, 8.5], dtype=f)
    ids, sc, _ = oracle.topk(rows, oracle.F32, oracle.L2, q, 5)
    known["simple"] = dict(file="simple_f32.mvf", space="embeddings", query=q.tolist(), k=5,
                           ids=[int(i) for i in ids], score_bits=bits(sc))
    # 2. examples/similarity_search.rs
    rows = clustered_rows()
    write("clustered_f32.mvf", fmt.build_mvf([dict(name="clustered_data", dimension=4, vector_type=0,
                                                    distance_metric=0, data_type=0, data=rows.tobytes())]))
    known["clustered"] = []
    for qq in ([1, 1, 1, 1], [5, 5, 5, 5], [-2, 3, 0, 4], [0, 0, 0, 0]):
        q = np.array(qq, dtype=f)
        ids, sc, _ = oracle.topk(rows, oracle.F32, oracle.L2, q, 5)
        ids_w, sc_w, _ = oracle.topk(rows, oracle.F32, oracle.L2, q, 5, mode=oracle.AS_WRITTEN)
        known["clustered"].append(dict(file="clustered_f32.mvf", space="clustered_data", query=q.tolist(),
                                       k=5, ids=[int(i) for i in ids], score_bits=bits(sc),
                                       as_written_ids=[int(i) for i in ids_w]))
    # 3. src/tests/test_utils.rs fixture
    rows = np.array([[1, 2, 3, 4], [5, 6, 7, 8], [9, 10, 11, 12]], dtype=f)
    write("test_space.mvf", fmt.build_mvf([dict(name="test_space", dimension=4, vector_type=0,
                                                 distance_metric=0, data_type=0, data=rows.tobytes())]))
    # 4. Float16 space as the reference builder writes it (total_vectors halved, builder.rs:476)
    vals = np.array([[3.14159265, 2.718281828, 1.0, -0.5]] * 2 + [[0.1 * i, -0.2 * i, 0.3 * i, 1.0] for i in range(6)],
                    dtype=f)
    h = np.array([[oracle.f32_to_f16(v) for v in r] for r in vals], dtype=np.uint16)
    write("f16_refbug.mvf", fmt.build_mvf([dict(name="half_space", dimension=4, vector_type=0,
                                                 distance_metric=2, data_type=1, data=h.tobytes())],
                                          reference_total_vectors_bug=True))
    # 5. several spaces + metadata columns + string heap
    a = oracle.synth(oracle.F32, 11, 0, 16, 8)
    b = oracle.synth(oracle.U8, 12, 0, 10, 16)
    c = oracle.synth(oracle.I8, 13, 0, 7, 16)
    d = oracle.synth(oracle.F16, 14, 0, 9, 12)
    write("multi.mvf", fmt.build_mvf(
        [dict(name="alpha", dimension=8, vector_type=0, distance_metric=2, data_type=0, data=a.tobytes()),
         dict(name="bytes_u8", dimension=16, vector_type=0, distance_metric=0, data_type=3, data=b.tobytes()),
         dict(name="bytes_i8", dimension=16, vector_type=0, distance_metric=1, data_type=2, data=c.tobytes()),
         dict(name="halfs", dimension=12, vector_type=0, distance_metric=2, data_type=1, data=d.tobytes())],
        metadata_columns=[dict(name="ids", data_type=4, data=struct.pack("<16I", *range(16))),
                          dict(name="labels", data_type=6, data=b"\x00" * 16)],
        string_heap=b"cat\x00dog\x00"))
    # 6. version 99 / bad footer fixtures are built on the fly in the tests (they are tiny)
    # 7. seeded search cases: DB is regenerated from (dtype, seed, n, d); expectations from the oracle
    cases = []
    spec = [(oracle.F32, 128, 3000), (oracle.F16, 96, 2500), (oracle.U8, 128, 4000), (oracle.I8, 64, 3500),
            (oracle.F32, 20, 777), (oracle.U8, 16, 1000)]
    seed = 1000
    for dtype, dd, n in spec:
        db = oracle.synth(dtype, seed, 0, n, dd)
        for metric in (oracle.L2, oracle.IP, oracle.COS):
            qs = oracle.synth(dtype, seed + 1, 0, 2, dd)
            qs[1] = db[n // 3]  # a query that is an exact copy of a row
            for qi in range(2):
                query = qs[qi].astype(np.float32) if dtype in (oracle.F32, oracle.F16) else qs[qi]
                k = 10 if qi == 0 else 33
                ids, sc, isc = oracle.topk(db, dtype, metric, query, k)
                cases.append(dict(dtype=dtype, d=dd, n=n, seed=seed, metric=metric, qseed=seed + 1, qrow=qi,
                                  planted=(n // 3 if qi == 1 else None), k=k, ids=[int(i) for i in ids],
                                  score_bits=bits(sc), iscores=[int(i) for i in isc]))
        seed += 10
    known["seeded"] = cases
    with open(os.path.join(OUT, "known_answers.json"), "w") as fh:
        json.dump(known, fh, indent=0, separators=(",", ":"))
    print("wrote fixtures to", OUT)


if __name__ == "__main__":
    main()
