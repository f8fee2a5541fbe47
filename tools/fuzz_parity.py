#!/usr/bin/env python
"""Random sweep over the test-stream writer's options: every stream goes through the reference (oracle) and through the CPU
emulation of the kernel bodies (tests/hostemu); any difference in error code, stride or pixels is printed.
usage: python tools/fuzz_parity.py <seed> <seconds>      (round 1: ~40 000 streams, no difference; round 2, with passes / local trees / wide
coefficient values / delta palettes among the options: see DESIGN.md §6)"""
import sys, random, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tools import streamgen
from tests import hostemu
from oracle import ref
rnd = random.Random(int(sys.argv[1]))
t0=time.time(); n=0; bad=0; genfail=0; referr=0
while time.time()-t0 < float(sys.argv[2]):
    kind = rnd.choice(["vardct","vardct","modular"])
    w = rnd.choice([8, 16, 24, 64, 100, 255, 256, 257, 300, 512, 520, 777, 1030]); h = rnd.choice([8, 16, 40, 64, 120, 256, 264, 400, 520])
    if kind == "vardct":
        kw = dict(mix=rnd.choice([0,1,1,2]), tree=rnd.choice([0,1,2]), ans=rnd.choice([0,1,1]), cfl=rnd.choice([0,1]),
                  smooth=rnd.choice([0,1]), extra_prec=rnd.choice([0,1,2,3]), orders=rnd.choice([0,0x1f,0x1fff,0xaa]),
                  alpha=rnd.choice([0,1]), raw_dq=rnd.choice([0,0,1,0x10,0x11]), permuted=rnd.choice([0,1]), container=rnd.choice([0,1]),
                  hfmul=rnd.choice([2,4,10,20,40]), hfmul_var=rnd.choice([0,4]), lz77=rnd.choice([0,0,1]), cfl_base=rnd.choice([0,1]),
                  x_qm=rnd.choice([0,2,3,5,7]), b_qm=rnd.choice([0,2,4,7]), las=rnd.choice([5,6,7,8]), clusters=rnd.choice([1,4,16,64]),
                  global_scale=rnd.choice([1000, 4000, 20000]), quant_lf=rnd.choice([4, 16, 64]), deadzone=rnd.choice([0.3,0.55,0.9]))
        if rnd.random() < 0.15: kw["force"] = rnd.randrange(27)
        ngroups = ((w+255)//256)*((h+255)//256)
        kw["presets"] = rnd.choice([1, min(2, ngroups), min(3, ngroups)])
        kw["block_ctx"] = rnd.choice([0,1]) if kw["ans"] else 0   # (the writer's block_ctx + prefix combination is broken)
        if kw["permuted"] and ngroups == 1: kw["permuted"] = 0
        if kw["container"]: kw["jxlp"] = rnd.choice([0,1])
        if not (kw["smooth"] and kw["x_qm"] == 3 and kw["b_qm"] == 2): kw["explicit_fh"] = 1
        else: kw["explicit_fh"] = rnd.choice([0,1])
        # round-2 features: passes, MA trees local to LF-group sections, coefficient values outside 16 bits, the transform mix
        if ngroups > 1 and rnd.random() < 0.4: kw["passes"] = rnd.choice([2, 3, 5])
        if rnd.random() < 0.3: kw["lf_local_tree"] = rnd.choice([1, 2, 3, 7])
        if rnd.random() < 0.25: kw["coef_spike"] = rnd.choice([32767, 32768, 40000, 1 << 21])
        if kw["mix"] == 1 and rnd.random() < 0.3: kw["big_take"] = rnd.choice([0.1, 0.35, 0.8]); kw["big_thr"] = rnd.choice([0.01, 0.03, 0.2])
        kw["seed"] = rnd.randrange(100000)
        try: d,_ = streamgen.vardct(w,h,**kw)
        except Exception as e: genfail+=1; continue
    else:
        kw = dict(tree=rnd.choice([0,1,2]), ans=rnd.choice([0,1]), lz77=rnd.choice([0,1]), alpha=rnd.choice([0,1]), palette=rnd.choice([0,0,1]),
                  local_tree=rnd.choice([0,0,1,2]), group_shift=rnd.choice([7,8,9,10]), rct=rnd.choice([-1]+list(range(0,42,5))), smooth=rnd.choice([0,1]),
                  clusters=rnd.choice([1,2,8,32]), container=rnd.choice([0,1]))
        if kw["palette"] and kw["tree"] == 2: kw["tree"] = 1
        if kw["palette"] and rnd.random() < 0.5: kw["pal_deltas"] = rnd.choice([3, 20, 50]); kw["pal_pred"] = rnd.choice([0, 1, 4, 5, 6, 13])
        kw["seed"] = rnd.randrange(100000)
        try: d,_ = streamgen.modular(w,h,**kw)
        except Exception as e: genfail+=1; continue
    a,ea,_,sa = ref.decode(d); b,eb,sb = hostemu.decode(d)
    n+=1
    if ea: referr+=1
    ok = ea==eb and ((a is None and b is None) or (a is not None and b is not None and sa==sb and np.array_equal(a,b)))
    if not ok:
        bad+=1; print("MISMATCH", kind, w, h, kw, repr(ea), repr(eb), flush=True)
print("cases", n, "mismatches", bad, "genfail", genfail, "ref errors", referr)
