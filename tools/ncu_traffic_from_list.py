#!/usr/bin/env python
"""profiles/ncu_traffic.json from an ncu launch list with DRAM metrics:
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file L.csv \
      python bench.py --steps 1 --warmup 3 --skip-e2e --skip-latency --frames-per-gpu 32 --streams 1
Sums, per kernel group of bench.py, the DRAM bytes of the launches of ONE decode (the last complete one in the list) and
divides by the frames of the batch. usage: python tools/ncu_traffic_from_list.py L.csv 32 [preset]"""
import csv, json, os, sys, collections
path, frames = sys.argv[1], int(sys.argv[2])
preset = sys.argv[3] if len(sys.argv) > 3 else "d1"
rows = list(csv.reader(open(path)))
hdr = next(r for r in rows if "Kernel Name" in r)
ik, im, iv, iu, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
launches = collections.OrderedDict()
for r in rows:
    if len(r) != len(hdr) or not r[iid].isdigit():
        continue
    d = launches.setdefault(int(r[iid]), {"name": r[ik]})
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1, "second": 1e3}.get(u, 1)
    d[r[im]] = v * scale
seq = list(launches.values())
# one decode = from a k_lf_chan run up to and including the next k_back_tile*; take the last complete one
ends = [i for i, l in enumerate(seq) if "k_back_tile" in l["name"]]
end = ends[-1]
start = ends[-2] + 1 if len(ends) > 1 else 0
GROUP = lambda n: ("k_lf_chan" if "k_lf_chan" in n else "k_hf_prep+k_hf_group" if "k_hf_" in n else "k_back_tile" if "k_back_tile" in n
                   else "k_back_generic" if "k_back_generic" in n else "k_lf_post+k_lf_llf" if "k_lf_" in n else "other")
out = {"_note": f"DRAM bytes per frame and kernel group, launches of one decode of {frames} 3840x2160 frames ({preset} preset), "
                "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum per launch (cold cache, serialised)", "preset": preset}
agg = collections.OrderedDict()
for l in seq[start:end + 1]:
    g = GROUP(l["name"])
    a = agg.setdefault(g, {"dram": 0.0, "ms": 0.0, "launches": 0})
    a["dram"] += l.get("dram__bytes_read.sum", 0) + l.get("dram__bytes_write.sum", 0)
    a["ms"] += l.get("gpu__time_duration.sum", 0)
    a["launches"] += 1
for g, a in agg.items():
    out[g] = {"dram_bytes_per_frame": a["dram"] / frames, "launches_per_decode": a["launches"], "ms_under_ncu": a["ms"], "preset": preset}
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
json.dump(out, open(os.path.join(root, "profiles", "ncu_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
