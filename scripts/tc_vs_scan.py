"""Large-size cross-check of the two GPU paths: the tensor path (K2) must return the same rows as the
HBM scan (K1) on a synthetic space too big for the CPU oracle.  Integer spaces: identical ids.
Float spaces: the score lists must agree within tolerance (accumulation order differs)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import metrovector_b200 as mv

dtype = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20_000_000
d = int(sys.argv[3]) if len(sys.argv) > 3 else 128
nq = int(sys.argv[4]) if len(sys.argv) > 4 else 256
k = int(sys.argv[5]) if len(sys.argv) > 5 else 10
eng = mv.GpuEngine(0)
gs = eng.synthetic(mv.DataType(dtype), mv.DistanceMetric.L2, n, d, 2024)
q = mv.synth_fill_host(dtype, 77, 0, nq, d)
rc = 0
for metric in (0, 1, 2):
    eng.set_path(2); ids_t, sc_t, _ = gs.search(q, k, metric=metric)
    eng.set_path(1); ids_s, sc_s, _ = gs.search(q, k, metric=metric)
    if dtype >= 2:
        bad = int(np.sum(np.any(ids_t != ids_s, axis=1)))
    else:
        bad = int(np.sum(~np.all(np.abs(sc_t - sc_s) <= 1e-4 * np.maximum(1.0, np.abs(sc_s)), axis=1)))
    print("dtype", dtype, "n", n, "d", d, "nq", nq, "k", k, "metric", metric, "queries that differ:", bad, flush=True)
    rc |= bad != 0
sys.exit(rc)
