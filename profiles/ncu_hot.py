#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export: hottest SASS instructions by executed count and by stall samples."""
import csv
import sys

path = sys.argv[1]
thresh = float(sys.argv[2]) if len(sys.argv) > 2 else 0.6
rows = list(csv.reader(open(path)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    try:
        data.append((int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]]), r[ix["Source"]].strip(), r))
    except Exception:
        pass
tot = sum(d[0] for d in data)
tots = sum(d[1] for d in data)
print("total warp instructions", tot, "samples", tots, "sass lines", len(data))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(d[3][ix[s]] or 0) for d in data) for s in stalls}
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
for i, (n, s, src, r) in enumerate(data):
    if n > tot * thresh / 100 or s > tots * thresh / 100:
        top = sorted(((int(r[ix[st]] or 0), st) for st in stalls), reverse=True)[:2]
        print("%4d exec %6.2f%% samp %6.2f%%  %-70s %s" % (i, 100 * n / tot, 100 * s / max(1, tots), src[:70],
                                                          " ".join("%s=%d" % (b, a) for a, b in top if a)))
