#!/usr/bin/env python
"""bench.py -- brute-force top-k QPS of the B200 scan, with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload D] [--impl ours|reference]

One process per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE for N>1).  A "step" is one search of
one query batch over the whole (row-sharded) database: kernels + merge (+ one NCCL allgather and
the global merge when N>1).  Rank 0 prints ONE JSON line.

  value  QPS with the query batch already resident in HBM, timed with CUDA events on the engine's
         stream, max over ranks.
  e2e    QPS through the public call a user makes (mvf_gpu_search with HOST buffers): per step the
         H2D copy of the queries from pinned memory, the search, the D2H copy of ids+scores, and
         the stream sync are all inside the timed region.
  roofline  the dominant kernel alone (CUDA events around it, on its stream) against the measured
         peaks in MEASURED_PEAKS.json; algorithmic bytes = local_rows * dimension * elem_size per
         launch on the bandwidth path, algorithmic flops = 2 * batch * local_rows * dimension on the
         tensor path.
  batched  the SAME resident database searched with the workload's large batch (D -> 1024 queries,
         E -> 1024, B -> 1024, C is batched already): the tensor-core path, same keys as the main
         leg (value, ms_per_step, e2e, roofline{bound: tensor}), all 128 queries of one query tile
         cross-checked against the bandwidth scan.  `value` stays the batch-1 figure.
  multirank_parity  (N>1) a small sharded search compared bit for bit with the same search on an
         unsharded copy held by rank 0, on both kernel paths, before anything is timed.
  cpu_baseline  the CPU oracle (a port of the reference's scan; the reference is Rust and cannot be
         built in this image) timed on this box's cores on a bounded row sample (>= 2 GiB of rows)
         scaled to N rows; `faithful_1thread` is the same scan the way the reference runs it (one
         thread, one Vec<f32> per row, examples/similarity_search.rs:147-169).

Workloads are BASELINE.json's configs; the database is synthetic (counter-based generator, filled
on the device), row-sharded over the ranks ("strong" scaling: the global database is fixed).
Default workload: D = 100M x 768 f16 cosine top-10, batch 1 -- the configuration north_star's target
is quoted on; its 153.6 GB fit one B200, so the same workload runs at every N.  `--workload B` is
BASELINE.json configs[1] (1M x 768 f32 L2 top-10, batch 1).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# name: (rows, dim, dtype, metric, k, batch)      dtype/metric codes = schema/types.fbs
WORKLOADS = {
    "A": (100_000, 128, 0, 2, 10, 1),            # examples/similarity_search-sized: 100k x 128 f32 cosine
    "B": (1_000_000, 768, 0, 0, 10, 1),          # 1M x 768 f32 L2 top-10, batch 1 (HBM-bound scan)
    "B1024": (1_000_000, 768, 0, 0, 10, 1024),   # the same space, batch 1024 (split-tf32 tensor path, K2')
    "C": (10_000_000, 768, 1, 2, 100, 4096),     # 10M x 768 f16 cosine top-100, batch 4096 (tensor path)
    "D": (100_000_000, 768, 1, 2, 10, 1),        # 100M x 768 f16 cosine top-10, batch 1
    "D1024": (100_000_000, 768, 1, 2, 10, 1024),
    "E": (1_000_000_000, 128, 3, 0, 10, 1),      # 1B x 128 u8 L2 top-10 (SIFT1B-shaped)
    "E1024": (1_000_000_000, 128, 3, 0, 10, 1024),
}
BATCHED_SIBLING = {"B": "B1024", "D": "D1024", "E": "E1024"}   # second timed leg on the same resident database
DT_NAME = {0: "f32", 1: "f16", 2: "i8", 3: "u8"}
ARITH = {0: "f32", 1: "f32", 2: "int32", 3: "int32"}
METRIC_NAME = {0: "L2", 1: "InnerProduct", 2: "Cosine"}
ES = {0: 4, 1: 2, 2: 1, 3: 1}
DB_SEED, Q_SEED = 0x4D564631, 0x4D564632
# DRAM traffic of the dominant kernel per launch, as bytes moved per algorithmic byte (rows read once), from the
# `ncu --set full` captures under profiles/ (dram__bytes_read.sum + dram__bytes_write.sum).  Keyed by workload; the
# value names the profile it came from so a stale entry is visible in the JSON line.
NCU_TRAFFIC = {
    "B": (3.076061e9 / 3.072e9, "profiles/r1_ncu_B.md"),
    "D": (154.014740e9 / 153.6e9, "profiles/r1_ncu_D.md"),
    "E": (25.604293e9 / 25.6e9, "profiles/r1_ncu_E.md"),
}
try:    # tensor-kernel entries are regenerated with the kernel (profiles/summarize.py writes this file)
    with open(os.path.join(ROOT, "profiles", "traffic_ratios.json")) as _fh:
        for _k, _v in json.load(_fh).items():
            NCU_TRAFFIC[_k] = (float(_v["ratio"]), _v["source"])
except Exception:
    pass


def describe(wl, n, d, dt, metric, k, batch):
    return "%s: %dx%d %s %s top-%d batch %d" % (wl, n, d, DT_NAME[dt], METRIC_NAME[metric], k, batch)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return {"hbm": float(p["hbm_gbs"]), "tensor": float(p["bf16_tflops"]),
                "tensor_sustained": float(p.get("bf16_tflops_sustained", 0.0)) or None, "source": "measured (MEASURED_PEAKS.json)"}
    except Exception:
        return {"hbm": 6650.0, "tensor": 1590.0, "tensor_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU while the measurement loops run."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        while not self.stop_flag and self.ok:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def finish(self):
        self.stop_flag = True
        self.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---- CPU arm (oracle port of the reference's scan) -----------------------------------------------
def cpu_sample_rows(n, d, dt, gib):
    return int(min(n, max(4096, int(gib * (1 << 30)) // (d * ES[dt]))))


def cpu_baseline_leg(n, d, dt, metric, k, gib=2.0):
    """All-core and faithful single-thread scans of a bounded sample of the workload; QPS scaled to n rows."""
    import oracle
    threads = os.cpu_count() or 1
    sample = cpu_sample_rows(n, d, dt, gib)
    db = oracle.synth(dt, DB_SEED, 0, sample, d, threads=threads)
    q = oracle.synth(dt, Q_SEED, 0, 4, d)
    if dt in (0, 1):
        q = q.astype(np.float32)
    oracle.scan_timed(db, dt, metric, q[:1], k, threads)            # warm the pages
    times = [oracle.scan_timed(db, dt, metric, q[i:i + 1], k, threads)[0] for i in range(4)]
    per_query_sample = float(np.median(times))
    qps = 1.0 / (per_query_sample * (n / sample))
    # the reference is single-threaded and materialises every row (as_f32): time that on a smaller slice
    fsample = int(min(sample, max(4096, (192 << 20) // (d * ES[dt]))))
    fsec, _, _ = oracle.scan_timed(db[:fsample], dt, metric, q[:1], k, 1, faithful=True)
    faithful_qps = 1.0 / (fsec * (n / fsample))
    return {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
            "sample": "%d of %d rows (%.2f GiB), median of 4 queries: %.3f s per query on the sample, scaled linearly to all rows"
                      % (sample, n, sample * d * ES[dt] / 2 ** 30, per_query_sample),
            "rows_timed": sample, "extrapolated": sample < n,
            "faithful_1thread": {"value": faithful_qps, "unit": "queries/s", "cores": 1, "rows_timed": fsample,
                                 "sample": "one thread, one Vec<f32> per row like the reference loop; %d rows in %.3f s, scaled linearly"
                                           % (fsample, fsec)}}


def run_reference(args, rank, world):
    """--impl reference: the reference's scan on the host CPU (oracle port, all host threads), rank 0 only.
    A step = one query over a bounded sample of the workload's rows (>= 2 GiB); ms_per_step is the measured time of
    that step, value is the workload's QPS obtained by scaling the median step linearly to all rows."""
    if rank != 0:
        return
    n, d, dt, metric, k, batch = WORKLOADS[args.workload]
    import oracle
    threads = os.cpu_count() or 1
    sample = cpu_sample_rows(n, d, dt, 2.0)
    db = oracle.synth(dt, DB_SEED, 0, sample, d, threads=threads)
    qs = oracle.synth(dt, Q_SEED, 0, args.warmup + args.steps, d)
    if dt in (0, 1):
        qs = qs.astype(np.float32)
    for i in range(args.warmup):
        oracle.scan_timed(db, dt, metric, qs[i:i + 1], k, threads)
    times = []
    for i in range(args.steps):
        t0 = time.perf_counter()
        oracle.scan_timed(db, dt, metric, qs[args.warmup + i:args.warmup + i + 1], k, threads)
        times.append(time.perf_counter() - t0)
    step_s = float(np.median(times))                                 # one query over the sample
    qps = 1.0 / (step_s * (n / sample))                               # a batch is `batch` independent scans on the CPU
    fsample = int(min(sample, max(4096, (192 << 20) // (d * ES[dt]))))
    fsec, _, _ = oracle.scan_timed(db[:fsample], dt, metric, qs[:1], k, 1, faithful=True)
    faithful = {"value": 1.0 / (fsec * (n / fsample)), "unit": "queries/s", "cores": 1, "rows_timed": fsample,
                "sample": "one thread, one Vec<f32> per row like the reference loop; %d rows in %.3f s, scaled linearly" % (fsample, fsec)}
    sample_txt = ("%d of %d rows (%.2f GiB) per step, median of %d steps = %.3f s, scaled linearly to all rows"
                  % (sample, n, sample * d * ES[dt] / 2 ** 30, args.steps, step_s))
    line = {
        "impl": "reference", "metric": "brute-force top-k QPS", "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": ARITH[dt], "data": "synthetic",
        "config": {"workload": describe(args.workload, n, d, dt, metric, k, batch), "rows_per_gpu": n // max(1, world),
                   "parallelism": "host CPU, %d threads (the reference itself is single-threaded)" % threads,
                   "l2_policy": "sample per step (%.2f GiB) exceeds every cache" % (sample * d * ES[dt] / 2 ** 30),
                   "path": "cpu"},
        "step_definition": "one query over %d of the %d rows; value = 1 / (median step x %d / %d)" % (sample, n, n, sample),
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port", "sample": sample_txt,
                         "rows_timed": sample, "extrapolated": sample < n, "faithful_1thread": faithful},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "ms_per_step_full_workload_extrapolated": step_s * (n / sample) * batch * 1e3,
    }
    print(json.dumps(line), flush=True)


# ---- GPU arm ---------------------------------------------------------------------------------------
def ids_agree(dt, ids_a, sc_a, ids_b, sc_b):
    """Two result sets of the same queries: integers bit-identical; floats by the id rule of tests/util.py --
    scores agree within tolerance and ids may differ only where the score gap to the k-th is inside it."""
    if dt >= 2:
        return bool(np.array_equal(ids_a, ids_b) and np.array_equal(sc_a, sc_b))
    if not np.allclose(sc_a, sc_b, rtol=1e-4, atol=2e-5):
        return False
    for q in range(ids_a.shape[0]):
        if np.array_equal(ids_a[q], ids_b[q]):
            continue
        tol = 2e-5 * max(1.0, abs(float(sc_a[q, -1])))
        kth = min(float(sc_a[q, -1]), float(sc_b[q, -1]))
        clear_a = set(ids_a[q][sc_a[q] < kth - 2 * tol].tolist())
        clear_b = set(ids_b[q][sc_b[q] < kth - 2 * tol].tolist())
        if not (clear_a <= set(ids_b[q].tolist()) and clear_b <= set(ids_a[q].tolist())):
            return False
    return True


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("MVF_BENCH_WORKLOAD", "D"), choices=sorted(WORKLOADS))
    ap.add_argument("--rows", type=int, default=0, help="override the row count (debug)")
    ap.add_argument("--path", type=int, default=0, help="0 auto, 1 scan kernel, 2 tensor-core kernel")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-batched", action="store_true", help="skip the batched (tensor-core) leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    lib = os.path.join(ROOT, "metrovector_b200", "libmvf_b200.so")
    if not os.path.exists(lib) and local_rank == 0:      # fresh checkout on the box: build in-tree (nvcc, sm_100a)
        import subprocess
        subprocess.run(["make", "-j8", "-C", os.path.join(ROOT, "metrovector_b200", "csrc")], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    for _ in range(600):                                    # other ranks wait for rank 0's build
        if os.path.exists(lib):
            break
        time.sleep(1)
    import metrovector_b200 as mv

    n, d, dt, metric, k, batch = WORKLOADS[args.workload]
    if args.rows:
        n = args.rows
    es = ES[dt]
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    stream = torch.cuda.Stream()
    eng = mv.GpuEngine(local_rank)
    eng.set_stream(stream.cuda_stream)
    eng.set_path(args.path)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(mv.GpuEngine.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        eng.comm_init(bytes(uid.cpu().numpy().tobytes()), rank, world)

    # ---- N>1: sharded == unsharded, bit for bit, on both kernel paths, before anything is timed ----------
    multirank_parity = None
    if world > 1:
        ok = True
        solo = mv.GpuEngine(local_rank) if rank == 0 else None
        for pdt, pmetric, pn, pd, pk, pnq in ((3, 0, 300_000, 128, 10, 3), (1, 2, 200_000, 96, 10, 2),
                                              (3, 0, 300_000, 128, 10, 64), (1, 2, 200_000, 96, 10, 64), (0, 0, 120_000, 64, 10, 48)):
            sh = eng.synthetic(pdt, pmetric, pn, pd, 99)
            pq = mv.synth_fill_host(pdt, 98, 0, pnq, pd)
            if pdt == 1 and pnq < 4:
                pq = pq.astype(np.float32)
            a = sh.search(pq, pk)
            sh.free()
            if rank == 0:
                whole = solo.synthetic(pdt, pmetric, pn, pd, 99)
                b = whole.search(pq, pk)
                whole.free()
                ok = ok and bool(np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
                                 and np.array_equal(a[2], b[2]))
        if solo is not None:
            solo.shutdown()
        multirank_parity = ok
        barrier()

    gs = eng.synthetic(dt, metric, n, d, DB_SEED)             # this rank's shard, generated in HBM

    def run_leg(batch, k, steps, warmup, cross_rows):
        """Times one (batch, k) on the resident space: device-resident QPS, e2e through host buffers, the dominant
        kernel alone; tensor-path legs are cross-checked against the scan kernel on `cross_rows` queries."""
        qnp = np.float32 if dt in (0, 1) else (np.int8 if dt == 2 else np.uint8)
        qdt = 0 if dt in (0, 1) else dt
        if dt == 1 and batch >= 4:
            qnp, qdt = np.float16, 1                          # the tensor path takes queries in the space's dtype
        nb = warmup + steps
        distinct = min(nb, max(1, (256 << 20) // max(1, batch * d * np.dtype(qnp).itemsize)))
        host_q = mv.synth_fill_host(dt, Q_SEED, 0, distinct * batch, d)
        if qnp == np.float32 and dt == 1:
            host_q = host_q.astype(np.float32)
        q_pinned = torch.from_numpy(np.ascontiguousarray(host_q.reshape(distinct, batch * d))).pin_memory()
        q_dev = q_pinned.cuda()
        qbytes = batch * d * np.dtype(qnp).itemsize
        ids_dev = torch.empty((batch, k), dtype=torch.int64, device="cuda")
        sc_dev = torch.empty((batch, k), dtype=torch.float32, device="cuda")
        ids_host = torch.empty((batch, k), dtype=torch.int64).pin_memory()
        sc_host = torch.empty((batch, k), dtype=torch.float32).pin_memory()
        torch.cuda.synchronize()

        def step_device(i):
            gs.search_device(q_dev.data_ptr() + (i % distinct) * qbytes, qdt, batch, k, metric, ids_dev.data_ptr(), sc_dev.data_ptr())

        def step_host(i):
            gs.search_raw(q_pinned.data_ptr() + (i % distinct) * qbytes, qdt, batch, k, metric, ids_host.data_ptr(), sc_host.data_ptr())

        sampler = ClockSampler(local_rank)
        sampler.start()
        with torch.cuda.stream(stream):
            # ---- device-resident throughput -----------------------------------------------------
            for i in range(warmup):
                step_device(i)
            barrier()
            eng.profile(False)                                     # resets the launch counters
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            for i in range(steps):
                step_device(warmup + i)
            ev1.record(stream)
            barrier()
            dev_ms = ev0.elapsed_time(ev1)
            launches = eng.profile_read()["all_launches"]
            # ---- end to end through the host-buffer call ----------------------------------------
            for i in range(warmup):
                step_host(i)
            barrier()
            t0 = time.perf_counter()
            for i in range(steps):
                step_host(warmup + i)
            torch.cuda.synchronize()
            e2e_s = time.perf_counter() - t0
            barrier()
            clocks = sampler.finish()
            # ---- the dominant kernel alone (roofline) -----------------------------------------------
            eng.profile(True)
            for i in range(steps):
                step_device(warmup + i)
            prof = eng.profile_read()
            eng.profile(False)
            # ---- tensor path only: one whole query tile of the last batch again through the bandwidth scan (K1) ----
            cross = None
            if prof["last_path"] == 2:
                nchk = min(batch, cross_rows)
                ids2 = torch.empty((nchk, k), dtype=torch.int64, device="cuda")
                sc2 = torch.empty((nchk, k), dtype=torch.float32, device="cuda")
                eng.set_path(1)
                jl = (warmup + steps - 1) % distinct
                gs.search_device(q_dev.data_ptr() + jl * qbytes, qdt, nchk, k, metric, ids2.data_ptr(), sc2.data_ptr())
                eng.set_path(args.path)
                torch.cuda.synchronize()
                cross = ids_agree(dt, ids_dev[:nchk].cpu().numpy(), sc_dev[:nchk].cpu().numpy(), ids2.cpu().numpy(), sc2.cpu().numpy())
        t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms = float(t[0]), float(t[1])

        # sanity of the last device result on rank 0: recompute the k returned rows' scores with the oracle
        verified = None
        if rank == 0:
            try:
                import oracle
                ids = ids_dev.cpu().numpy().astype(np.uint64)[0]
                sc = sc_dev.cpu().numpy()[0]
                j = (warmup + steps - 1) % distinct
                q0 = host_q.reshape(distinct, batch, d)[j, 0]
                ok = True
                for gid, s in zip(ids[:k], sc[:k]):
                    row = mv.synth_fill_host(dt, DB_SEED, int(gid), 1, d)[0]
                    ref, iref, _ = oracle.score_row(row, dt, metric, q0.astype(np.float32) if dt in (0, 1) else q0)
                    ok = ok and (abs(ref - s) <= 1e-3 * max(1.0, abs(ref)))
                verified = bool(ok and np.all(np.diff(sc[:k] if metric != 1 else -sc[:k]) >= 0) and cross is not False)
            except Exception as ex:   # the oracle is optional equipment for the bench
                verified = "unchecked: %s" % ex

        peaks = measured_peaks()
        scan_ms = prof["scan_ms_total"] / max(1, prof["scan_launches"])
        bytes_per_launch = gs.local_n * d * es
        tensor = prof["last_path"] == 2
        achieved = bytes_per_launch / (scan_ms * 1e-3) / 1e9 if scan_ms > 0 else 0.0
        wl_key = args.workload if batch == WORKLOADS[args.workload][5] else BATCHED_SIBLING.get(args.workload, "")
        ratio, ratio_src = NCU_TRAFFIC.get(wl_key, (None, None))
        roof = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm"], "unit": "GB/s", "frac": achieved / peaks["hbm"],
                "traffic": (bytes_per_launch * ratio) if ratio else None,
                "traffic_source": ("ncu --set full of this workload, %s, scaled by rows" % ratio_src) if ratio else None,
                "peak_source": peaks["source"], "kernel": "gemm_topk_tc" if tensor else "scan_topk",
                "kernel_ms": scan_ms, "kernel_launches_per_step": prof["scan_launches"] / max(1, steps),
                "algorithmic_bytes_per_launch": bytes_per_launch,
                # the measured peak is a COPY (half reads, half writes); a read-only scan can exceed it, so the
                # datasheet figure is reported next to it: 7.7 TB/s for HGX B200 (B200_PROFILING.md)
                "peak_nominal": 7700.0, "frac_of_nominal": achieved / 7700.0}
        if tensor:
            flops = 2.0 * batch * gs.local_n * d
            tf = flops / (scan_ms * 1e-3) / 1e12 if scan_ms > 0 else 0.0
            tpeak, tsust = peaks["tensor"], peaks["tensor_sustained"]
            # Which measured figure applies: the kernel is timed inside a run of back-to-back steps, and when the
            # sampler saw the power cap during that run the sustained cuBLAS figure is the ceiling that describes the
            # regime; otherwise the burst one.  Both are reported.  MEASURED_PEAKS.json holds a bf16 figure only: the int8
            # tensor rate is twice the 16-bit one and the tf32 rate half of it (datasheet ratios); the split-tf32 path
            # issues three tf32 MMAs per product, so its algorithmic flops are measured against a third of the tf32 rate.
            capped = "sw_power_cap" in clocks["reasons"] and tsust
            use = tsust if capped else tpeak
            scale = 2.0 if dt in (2, 3) else (1.0 / 6.0 if dt == 0 else 1.0)
            src = "%s bf16 figure of MEASURED_PEAKS.json%s" % (
                "sustained (run was power-capped)" if capped else "burst",
                " x2 for int8 operands" if dt in (2, 3) else (" / 2 for tf32 operands / 3 MMAs per product (split tf32)" if dt == 0 else ""))
            roof.update({"bound": "tensor", "achieved": tf, "peak": use * scale, "unit": "TOP/s" if dt in (2, 3) else "TFLOP/s",
                         "frac": tf / (use * scale), "peak_source": src, "algorithmic_flops_per_launch": flops,
                         "peak_burst": tpeak * scale, "peak_sustained": (tsust * scale) if tsust else None,
                         "frac_of_burst": tf / (tpeak * scale),
                         "peak_nominal": 2250.0 * scale, "frac_of_nominal": tf / (2250.0 * scale),   # dense f16 datasheet figure
                         "hbm_gbs": achieved})
        return {"value": steps * batch / (dev_ms * 1e-3), "ms_per_step": dev_ms / steps, "steps": steps, "warmup": warmup,
                "e2e": {"value": steps * batch / (e2e_ms * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": qbytes,
                        "d2h_bytes_per_step": batch * k * 12, "ms_per_step": e2e_ms / steps},
                "gpu_launches": int(launches), "roofline": roof, "clocks": clocks, "verified_against_oracle": verified,
                "cross_checked_vs_scan": cross, "path": "tensor" if tensor else "scan"}

    main_leg = run_leg(batch, k, args.steps, args.warmup, 128)
    batched = None
    sib = BATCHED_SIBLING.get(args.workload)
    if sib and not args.no_batched and args.path != 1:
        _, _, _, _, bk, bbatch = WORKLOADS[sib]
        bsteps = max(3, min(args.steps, 10))
        batched = run_leg(bbatch, bk, bsteps, 3, 128)
        batched["workload"] = describe(sib, n, d, dt, metric, bk, bbatch)
        batched["unit"] = "queries/s"

    xinfo = eng.exchange_info()
    if rank == 0:
        line = {
            "metric": "brute-force top-k QPS", "value": main_leg["value"], "unit": "queries/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_leg["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": ARITH[dt], "data": "synthetic",
            "config": {"workload": describe(args.workload, n, d, dt, metric, k, batch), "rows_per_gpu": gs.local_n,
                       "parallelism": ("row-sharded x%d, one exchange of batch*k candidates per rank (%s)" % (
                           world, "peer-memory kernel for <= 512 candidates, ncclAllGather above" if xinfo["peer_ready"]
                           else "ncclAllGather; peer path off: %s" % xinfo["note"])) if world > 1 else "1 GPU",
                       "l2_policy": "database per GPU (%.2f GB) %s L2 (126 MB)" %
                                    (gs.local_n * d * es / 1e9, "exceeds" if gs.local_n * d * es > 126e6 else "FITS IN"),
                       "path": main_leg["path"]},
            "e2e": main_leg["e2e"],
            "gpu_launches": main_leg["gpu_launches"] + (batched["gpu_launches"] if batched else 0),
            "roofline": main_leg["roofline"],
            "clocks": main_leg["clocks"],
            "verified_against_oracle": main_leg["verified_against_oracle"],
            "tensor_path_cross_checked_vs_scan": main_leg["cross_checked_vs_scan"],
        }
        if batched:
            line["batched"] = batched
        if world > 1:
            line["multirank_parity"] = multirank_parity
            line["exchange"] = {"peer_searches": xinfo["peer_searches"], "nccl_searches": xinfo["nccl_searches"]}
        if not args.no_cpu:
            try:
                line["cpu_baseline"] = cpu_baseline_leg(n, d, dt, metric, k)
            except Exception as ex:
                line["cpu_baseline"] = {"value": None, "unit": "queries/s", "cores": 0, "kind": "port", "sample": "failed: %s" % ex}
        print(json.dumps(line), flush=True)
    gs.free()
    eng.shutdown()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
