#!/usr/bin/env python
"""Aggregates the ncu source page by CUDA source line.

usage: python tools/ncu_hot_lines.py report.ncu-rep [topN] [function-substring]

Prints, for the functions whose name contains the substring (default: all), the share of warp-stall
samples and executed warp instructions per file:line, plus per-file totals and line-range buckets.
"""
import csv, subprocess, sys, collections, os

path = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
fsub = sys.argv[3] if len(sys.argv) > 3 else ""
txt = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
cur_file = cur_fn = ""
ci = cs = None
data = collections.defaultdict(lambda: [0.0, 0.0, ""])
for r in csv.reader(txt.splitlines()):
    if not r: continue
    if r[0] == "File Path": cur_file = os.path.basename(r[1]); continue
    if r[0] == "Function Name": cur_fn = r[1]; continue
    if r[0] == "Line No":
        ci = r.index("Instructions Executed"); cs = r.index("# Samples"); continue
    if ci is None or len(r) <= max(ci, cs) or not r[0].strip().isdigit(): continue
    if fsub and fsub not in cur_fn: continue
    try: inst = float(r[ci] or 0); samp = float(r[cs] or 0)
    except ValueError: continue
    d = data[(cur_file, int(r[0]))]
    d[0] += inst; d[1] += samp; d[2] = r[1].strip()[:110]
ti = sum(d[0] for d in data.values()) or 1
ts = sum(d[1] for d in data.values()) or 1
print(f"functions matching '{fsub}': warp-instructions {ti:.0f}, samples {ts:.0f}")
files = collections.defaultdict(lambda: [0.0, 0.0])
for (f, ln), d in data.items():
    files[f][0] += d[0]; files[f][1] += d[1]
for f, d in sorted(files.items(), key=lambda kv: -kv[1][1]):
    print(f"  file {f:24s} {100*d[1]/ts:5.1f}% samp {100*d[0]/ti:5.1f}% inst")
for (f, ln), d in sorted(data.items(), key=lambda kv: -kv[1][1])[:topn]:
    print(f"{100*d[1]/ts:5.1f}% samp {100*d[0]/ti:5.1f}% inst  {f}:{ln}: {d[2]}")
