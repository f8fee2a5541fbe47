#!/usr/bin/env python
"""Writes a corpus of corrupted test streams (several bit flips / random bytes in the header area, flips and random runs
anywhere) for tools/asan_sweep.sh.   usage: python tools/gen_corrupt_corpus.py <seed> <count> <out-dir>"""
import sys, random
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools import streamgen
rnd=random.Random(int(sys.argv[1])); n=0
os.makedirs(sys.argv[3], exist_ok=True)
while n < int(sys.argv[2]):
    kind=rnd.choice(["vardct","modular"])
    w=rnd.choice([8,64,100,200,257,300]); h=rnd.choice([8,64,120,136,264])
    if kind=="vardct":
        kw=dict(mix=rnd.choice([0,1,2]),tree=rnd.choice([0,1,2]),ans=rnd.choice([0,1]),alpha=rnd.choice([0,0,1]),raw_dq=rnd.choice([0,0,0x11]),
                container=rnd.choice([0,1]),lz77=rnd.choice([0,0,1]),orders=rnd.choice([0,0x1f]),block_ctx=0,seed=rnd.randrange(1000),
                passes=rnd.choice([1,1,2,3]),lf_local_tree=rnd.choice([0,0,1,2,3,7]),coef_spike=rnd.choice([0,0,0,40000,1<<21]))
        if w <= 256 and h <= 256: kw["passes"]=1
        if kw["container"]: kw["jxlp"]=rnd.choice([0,1])
        try: data=streamgen.vardct(w,h,**kw)[0]
        except Exception: continue
    else:
        kw=dict(tree=rnd.choice([0,1,2]),ans=rnd.choice([0,1]),lz77=rnd.choice([0,1]),alpha=rnd.choice([0,1]),palette=rnd.choice([0,0,1]),local_tree=rnd.choice([0,1,2]),
                group_shift=rnd.choice([7,8,9]),seed=rnd.randrange(1000),container=rnd.choice([0,1]))
        if kw["palette"]: kw["pal_deltas"]=rnd.choice([0,3,50]); kw["pal_pred"]=rnd.choice([1,4,5,6,13])
        if kw["palette"] and kw["tree"]==2: kw["tree"]=1
        try: data=streamgen.modular(w,h,**kw)[0]
        except Exception: continue
    for k in range(8):
        b=bytearray(data)
        style=k%4
        if style==0:   # several flips in the header area
            for _ in range(rnd.randrange(1,4)): b[rnd.randrange(0,min(len(b),220))]^=1<<rnd.randrange(8)
        elif style==1: # random bytes in the header area
            for _ in range(rnd.randrange(1,3)): b[rnd.randrange(0,min(len(b),400))]=rnd.randrange(256)
        elif style==2: # several flips anywhere
            for _ in range(rnd.randrange(2,6)): b[rnd.randrange(len(b))]^=1<<rnd.randrange(8)
        else:          # a run of random bytes
            p=rnd.randrange(len(b)); 
            for i in range(p,min(len(b),p+rnd.randrange(1,16))): b[i]=rnd.randrange(256)
        open(os.path.join(sys.argv[3], "c%s_%05d.jxl" % (sys.argv[1], n)), "wb").write(bytes(b)); n+=1
