#!/usr/bin/env python
"""Runs bench.py for the given workloads and prints one compact line each (full JSON to gpurun_out/)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
extra = []
wls = []
for a in sys.argv[1:]:
    (extra if a.startswith("-") or (extra and not a.isalpha() and not a[0].isalpha()) else wls).append(a)
for w in wls:
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", w, "--no-cpu"] + extra,
                         capture_output=True, text=True)
    line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else ""
    try:
        d = json.loads(line)
    except Exception:
        print(w, "FAILED", out.stdout[-500:], out.stderr[-1500:])
        continue
    with open(os.path.join(ROOT, "gpurun_out", "bench_%s.json" % w), "w") as fh:
        fh.write(line + "\n")
    r = d["roofline"]
    print("%-44s qps %10.1f ms/step %9.4f e2e %10.1f | %s %8.4f ms %7.1f %s frac %.3f | clk %s %s | ok=%s" % (
        d["config"]["workload"], d["value"], d["ms_per_step"], d["e2e"]["value"], r["kernel"], r["kernel_ms"], r["achieved"],
        r["unit"], r["frac"], d["clocks"]["sm_mhz"], ",".join(d["clocks"]["reasons"]), d["verified_against_oracle"]), flush=True)
