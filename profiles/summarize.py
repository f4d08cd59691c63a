#!/usr/bin/env python
"""Condenses the ncu exports in gpurun_out/ into profiles/<round>_*.md (what the judge reads) and, for the tensor kernel,
profiles/traffic_ratios.json (DRAM bytes per algorithmic byte, read by bench.py for roofline.traffic).

    python profiles/summarize.py [src_dir] [out_dir] [round_prefix, default r2] [tag=algorithmic_bytes ...]
"""
import csv
import glob
import os
import sys

src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
out = sys.argv[2] if len(sys.argv) > 2 else "profiles"
RND = sys.argv[3] if len(sys.argv) > 3 else "r2"
ALGO = dict((a.split("=")[0], float(a.split("=")[1])) for a in sys.argv[4:] if "=" in a)   # tag -> algorithmic bytes of the profiled launch
ratios = {}
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.avg.per_second", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__shared_mem_per_block_dynamic"]
for path in sorted(glob.glob(os.path.join(src, "prof_*_raw.csv"))):
    tag = os.path.basename(path)[5:-8]
    rows = list(csv.reader(open(path)))
    d, u = dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))
    lines = ["# ncu --set full, kernel `%s` (%s)" % (d.get("Kernel Name", "?"), tag), "",
             "grid %s block %s" % (d.get("Grid Size"), d.get("Block Size")), "", "| metric | value | unit |", "|---|---|---|"]
    for k in KEYS:
        if k in d:
            lines.append("| %s | %s | %s |" % (k, d[k], u[k]))
    # opcode mix + stall summary from the source page
    sp = path.replace("_raw.csv", "_source.csv")
    if os.path.exists(sp):
        import collections
        rr = list(csv.reader(open(sp)))
        hdr = rr[1]
        ix = {h: i for i, h in enumerate(hdr)}
        ops, tot, samples = collections.Counter(), 0, collections.Counter()
        for r in rr[2:]:
            try:
                n = int(r[ix["Instructions Executed"]])
            except Exception:
                continue
            op = r[ix["Source"]].strip().split()
            if op and op[0].startswith("@"):
                op = op[1:]
            ops[op[0].split(".")[0] if op else "?"] += n
            tot += n
            for h in hdr:
                if h.startswith("stall_") and "Not Issued" not in h:
                    samples[h] += int(r[ix[h]] or 0)
        lines += ["", "SASS lines: %d, warp instructions executed: %d" % (len(rr) - 2, tot), "",
                  "opcode mix: " + ", ".join("%s %.1f%%" % (k, 100.0 * v / max(1, tot)) for k, v in ops.most_common(14)), "",
                  "stall samples: " + ", ".join("%s %d" % (k, v) for k, v in samples.most_common(8) if v)]
    if tag in ALGO:
        def gb(key):
            v = float(d[key].replace(",", "")); unit = u[key]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}[unit]
        traffic = gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
        ratios[tag] = {"ratio": traffic / ALGO[tag], "source": "profiles/%s_ncu_%s.md" % (RND, tag), "dram_bytes": traffic,
                       "algorithmic_bytes": ALGO[tag]}
        lines += ["", "DRAM traffic %.4g B for %.4g algorithmic bytes of rows: x%.3f" % (traffic, ALGO[tag], traffic / ALGO[tag])]
    open(os.path.join(out, "%s_ncu_%s.md" % (RND, tag)), "w").write("\n".join(lines) + "\n")
    print("wrote", tag)
for path in glob.glob(os.path.join(src, "launches_*.csv")):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = {}
    for r in rows:
        agg.setdefault(r[4], []).append(float(r[-1]))
    tag = os.path.basename(path)[9:-4]
    tot = sum(sum(v) for v in agg.values())
    lines = ["# launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`), %s" % tag, "",
             "per-launch times are cold-cache and serialised: compare shares", "", "| kernel | launches | mean us | share |", "|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        lines.append("| %s | %d | %.1f | %.1f%% |" % (k, len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
    open(os.path.join(out, "%s_launches_%s.md" % (RND, tag)), "w").write("\n".join(lines) + "\n")
    print("wrote launches", tag)
if ratios:
    import json
    path = os.path.join(out, "traffic_ratios.json")
    old = {}
    if os.path.exists(path):
        old = json.load(open(path))
    old.update(ratios)
    json.dump(old, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path)
