#!/usr/bin/env python
"""BASELINE.json configs[3]: 8192x8192 lossless modular frame (fjxl-shaped: prefix codes + LZ77, YCoCg RCT, global MA
tree, 1024 groups) on one B200 -- decodes it through the batch API, checks every byte against the oracle (the
unmodified reference) and prints device / end-to-end times next to the reference's single-thread time.
usage (GPU box): python tools/c4_check.py [size]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import j40_b200 as J
from oracle import ref
from tools import streamgen

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
t = time.time(); data, st = streamgen.modular(n, n, seed=3); t_gen = time.time() - t
t = time.time(); want, e0, _, _ = ref.decode(data); t_ref = time.time() - t
assert e0 == ""
b = J.Batch(0)
b.add(data); b.upload()
ms = []
for _ in range(4):
    b.decode(); assert b.wait() == 0
    ms.append(b.last_decode_ms())
got = b.read_pixels(0)
ok = bool(np.array_equal(got, want))
km = b.kernel_ms()
t = time.time(); px, err, _, _ = J.decode(data); t_e2e = time.time() - t
ok2 = err == "" and bool(np.array_equal(px, want))
src = (streamgen.synth(n, n, 3) >> 2) << 2
lossless = bool(np.array_equal(got[..., :3], src))
print(json.dumps({"config": f"{n}x{n} modular lossless (prefix+LZ77, RCT 6, {st['sections']} sections)", "bytes": len(data),
                  "bit_exact_vs_reference": ok and ok2, "lossless_roundtrip": lossless,
                  "device_ms": min(ms), "device_mpix_s": n * n / min(ms) / 1e3, "kernel_ms": km,
                  "j40_api_end_to_end_s": t_e2e, "reference_single_thread_s": t_ref, "reference_mpix_s": n * n / t_ref / 1e6}))
