import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import metrovector_b200 as mv, oracle
import sys as _s
n, d, dtype = 30000, 128, int(_s.argv[1]) if len(_s.argv) > 1 else 3
db = mv.synth_fill_host(dtype, 41, 0, n, d)
eng = mv.GpuEngine(0)
gs = eng.from_host(db, mv.DataType(dtype), mv.DistanceMetric.L2)
eng.set_path(2)
q = mv.synth_fill_host(dtype, 42, 0, 200, d); q[0] = db[n // 2]
for rep in range(3):
    for metric in (0, 1):
        ids, sc, isc = gs.search(q, 10, metric=metric)
        bad = 0
        for qi in range(200):
            o_ids, _, _ = oracle.topk(db, dtype, metric, q[qi], 10)
            if ids[qi].tolist() != o_ids.tolist():
                bad += 1
                if bad <= 2:
                    missing = set(o_ids.tolist()) - set(ids[qi].tolist())
                    print("  q", qi, "missing", sorted(missing), "tiles", [m // 256 for m in missing], "cols", [m % 256 for m in missing])
        print("rep", rep, "metric", metric, "bad queries", bad, flush=True)
