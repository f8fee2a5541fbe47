#!/usr/bin/env python
"""Opcode mix of one kernel from an ncu report, weighted by executed warp instructions.
usage: python tools/ncu_sass_mix.py report.ncu-rep kernel-substring [units]   (units: divide counts, e.g. symbols)"""
import csv, subprocess, sys, collections
path, sub = sys.argv[1], sys.argv[2]
units = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
txt = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
cur = None; hdr = None; seen = set()
ops = collections.Counter(); samp = collections.Counter(); tot = 0.0
for r in csv.reader(txt.splitlines()):
    if not r: continue
    if r[0] == "Kernel Name":
        cur = r[1] if (sub in r[1] and r[1] not in seen) else None
        if cur: seen.add(cur)
        continue
    if r[0] == "Address": hdr = r; ie = r.index("Instructions Executed"); isamp = r.index("# Samples"); continue
    if not cur or hdr is None or len(r) < len(hdr) - 2: continue
    src = r[1].split()
    if not src: continue
    op = src[1] if src[0].startswith("@") and len(src) > 1 else src[0]
    op = op.split(".")[0]
    try: c = float(r[ie] or 0); s = float(r[isamp] or 0)
    except ValueError: continue
    ops[op] += c; samp[op] += s; tot += c
ts = sum(samp.values()) or 1
print(f"kernel '{sub}': {tot:.0f} warp instructions, {tot/units:.1f} per unit")
for op, c in ops.most_common(28):
    print(f"  {op:10s} {c/units:9.1f}  {100*c/tot:5.1f}% inst  {100*samp[op]/ts:5.1f}% samples")
