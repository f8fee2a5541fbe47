#!/usr/bin/env python
"""Aggregates the ncu source page by CUDA source line: python tools/ncu_hot_lines.py report.ncu-rep [topN]"""
import csv, subprocess, sys
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]
ci = hdr.index("Instructions Executed"); cs = hdr.index("# Samples")
data = []
for r in rows[hi + 1:]:
    if len(r) <= max(ci, cs) or not r[0].strip().isdigit(): continue   # only the per-source-line summary rows
    try: inst = float(r[ci] or 0); samp = float(r[cs] or 0)
    except ValueError: continue
    data.append((inst, samp, r[0], r[1].strip()[:120]))
ti = sum(d[0] for d in data) or 1; ts = sum(d[1] for d in data) or 1
print(f"total warp-instructions {ti:.0f}, samples {ts:.0f}")
for inst, samp, ln, src in sorted(data, key=lambda d: -d[1])[:topn]:
    print(f"{100*samp/ts:5.1f}% samp {100*inst/ti:5.1f}% inst  L{ln:>5s}: {src}")
