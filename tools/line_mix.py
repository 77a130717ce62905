#!/usr/bin/env python
"""Per-source-line instruction counts from `ncu --page source --csv --print-source cuda,sass`.
usage: ncu -i rep --page source --csv --print-source cuda,sass --launch-skip K --launch-count 1 | python tools/line_mix.py [top]"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
top = int(sys.argv[1]) if len(sys.argv) > 1 else 30
cur_file = None
agg = {}
ie = st = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        ie = r.index("Instructions Executed"); st = r.index("# Samples")
        continue
    if ie is None or len(r) <= ie:
        continue
    if r[0] not in ("", "Line No") and r[0].isdigit():
        try:
            n = int(r[ie]); s = int(r[st])
        except ValueError:
            continue
        key = (cur_file, int(r[0]))
        a = agg.setdefault(key, [0, 0, r[1].strip()])
        a[0] += n; a[1] += s
tot = sum(a[0] for a in agg.values()) or 1
print("total warp instr (line-attributed):", tot)
for (f, ln), (n, s, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% %11d exec %6d smp  %s:%d  %s" % (100.0 * n / tot, n, s, f, ln, src[:80]))
