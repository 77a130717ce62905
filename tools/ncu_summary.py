#!/usr/bin/env python
"""Summarise gpurun_out/{launches.csv,prof_*.ncu-rep} into profiles/<tag>_ncu.{json,md} (run here, no GPU).
usage: python tools/ncu_summary.py r01 [gpurun_out/prof_r01.ncu-rep]"""
import collections
import csv
import json
import subprocess
import sys

tag = sys.argv[1]
rep = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/prof_%s.ncu-rep" % tag
out = {"launch_list": [], "kernels": {}}

# 1. launch list: share of the step per kernel (cold-cache, serialised: compare SHARES)
try:
    rows = list(csv.reader(l for l in open("gpurun_out/launches.csv") if l.startswith('"')))
    hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
    agg = collections.defaultdict(list)
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = {"ns": v / 1000.0, "us": v, "ms": v * 1000.0, "s": v * 1e6}.get(r[ui], v)
        agg[r[ki]].append(v)
    mine = {k: v for k, v in agg.items() if "vdet::" in k}
    tot = sum(sum(v) for v in mine.values())
    for k, v in sorted(mine.items(), key=lambda kv: -sum(kv[1])):
        out["launch_list"].append({"kernel": k.split("(")[0], "launches": len(v), "avg_us": sum(v) / len(v),
                                   "share_of_vdet_time": sum(v) / tot})
except FileNotFoundError:
    pass

# 2. full capture: one row per captured launch
want = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__waves_per_multiprocessor": "waves",
}
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")].split("(")[0]
    d = {}
    for m, short in want.items():
        if m in hdr:
            i = hdr.index(m)
            d[short] = float(r[i].replace(",", "")) if r[i] not in ("", "n/a") else None
            d[short + "_unit"] = units[i]
    out["kernels"].setdefault(name, []).append(d)

json.dump(out, open("profiles/%s_ncu.json" % tag, "w"), indent=1)
with open("profiles/%s_ncu.md" % tag, "w") as f:
    f.write("# ncu summary %s (B200, --clock-control none)\n\n## launch list (gpu__time_duration, cold cache, serialised)\n\n" % tag)
    f.write("| kernel | launches | avg us | share of vdet time |\n|---|---|---|---|\n")
    for e in out["launch_list"]:
        f.write("| %s | %d | %.1f | %.1f %% |\n" % (e["kernel"], e["launches"], e["avg_us"], 100 * e["share_of_vdet_time"]))
    f.write("\n## --set full captures (first launch of each kernel)\n\n")
    for k, lst in out["kernels"].items():
        d = lst[0]
        f.write("### %s\n\n" % k)
        for short in want.values():
            if short in d and d[short] is not None:
                f.write("* %s: %s %s\n" % (short, d[short], d.get(short + "_unit", "")))
        f.write("\n")
print(open("profiles/%s_ncu.md" % tag).read())
