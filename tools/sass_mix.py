#!/usr/bin/env python
"""Instruction mix / hottest instructions of one kernel from `ncu --page source --csv --print-source sass`.
usage: ncu -i rep --page source --csv --print-source sass [--kernel-name regex:..] | python tools/sass_mix.py [top]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(sys.stdin))
top = int(sys.argv[1]) if len(sys.argv) > 1 else 25
kern = None
hdr = None
ops = collections.Counter(); samp = collections.Counter(); tot = 0; hot = []
def flush():
    global ops, samp, tot, hot
    if kern is None or tot == 0: return
    print("==", kern[:100], "total warp instr", tot)
    for op, n in ops.most_common(top):
        print("  %-10s %12d %5.1f%%  stall samples %6d" % (op, n, 100.0 * n / tot, samp[op]))
    print("  -- hottest by stall samples")
    for s, n, src in sorted(hot, reverse=True)[:top]:
        print("  %6d samples %10d exec  %s" % (s, n, src.strip()[:90]))
    ops = collections.Counter(); samp = collections.Counter(); tot = 0; hot = []
for r in rows:
    if not r: continue
    if r[0] == "Kernel Name":
        flush(); kern = r[1]; hdr = None; continue
    if r[0] == "Address":
        hdr = r; ie = r.index("Instructions Executed"); so = r.index("Source"); st = r.index("# Samples"); continue
    if hdr is None or len(r) <= ie: continue
    try: n = int(r[ie]); s = int(r[st])
    except ValueError: continue
    tot += n
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[so])
    op = m.group(2) if m else r[so][:10]
    ops[op] += n; samp[op] += s; hot.append((s, n, r[so]))
flush()
