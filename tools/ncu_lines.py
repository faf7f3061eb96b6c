#!/usr/bin/env python
"""Per-source-line share of executed warp instructions and stall samples from
`ncu -i rep --page source --csv --print-source cuda,sass > src.csv`.  usage: ncu_lines.py src.csv [top]"""
import collections
import csv
import os
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, hdr = None, None
inst, smp = collections.Counter(), collections.Counter()
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1]
        continue
    if len(r) > 10 and r[0] == "Line No":
        hdr = r
        continue
    if len(r) > 10 and hdr:
        try:
            line = int(r[0])
        except ValueError:
            continue

        def num(k):
            try:
                return int(r[hdr.index(k)])
            except (ValueError, IndexError):
                return 0
        inst[(cur, line)] += num("Instructions Executed")
        smp[(cur, line)] += num("# Samples")
ti, ts = sum(inst.values()), sum(smp.values())
print("total warp instructions %d, samples %d" % (ti, ts))
cache = {}
for (f, l), v in sorted(inst.items(), key=lambda kv: -kv[1])[:top]:
    if f not in cache:
        cache[f] = open(f).read().split("\n") if os.path.exists(f) else []
    text = cache[f][l - 1].strip()[:100] if 0 < l <= len(cache[f]) else ""
    print("%5.1f%% inst %5.1f%% smp  %s:%d  %s" % (100.0 * v / ti, 100.0 * smp[(f, l)] / max(ts, 1), os.path.basename(f), l, text))
