"""Dev tool: where does a step go on the strain workload?  usage: python tools/strain_probe.py [n_strains] [bases] [k]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sibelia_b200 as sb
from sibelia_b200 import synth

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 2
bl = int(float(sys.argv[2])) if len(sys.argv) > 2 else 125_000_000
k = int(sys.argv[3]) if len(sys.argv) > 3 else 25
chrs = synth.strains(ns, bl)
ctx = sb.Context(0)
ctx.upload(chrs)
for it in range(4):
    ctx.set_profiling(it == 3)
    t0 = time.perf_counter()
    count, ninst = ctx.enumerate_resident(k)
    wall = (time.perf_counter() - t0) * 1e3
    print("resident: V=%d I=%d device %.3f ms wall %.3f ms launches %d" % (count, ninst, ctx.last_device_ms(), wall, ctx.last_launches()))
tot = 0.0
for s in sorted(ctx.kernel_stats(), key=lambda s: -s["ms"]):
    print("   %-24s %8.3f ms x%d" % (s["name"], s["ms"], s["launches"]))
    tot += s["ms"]
print("   kernels total %.3f ms" % tot)
ctx.set_profiling(False)
for it in range(3):
    t0 = time.perf_counter()
    pos, neg = ctx.download()
    print("download: %.3f ms (%d bytes)" % ((time.perf_counter() - t0) * 1e3, pos.nbytes + neg.nbytes))
for it in range(3):
    t0 = time.perf_counter()
    c, pos, neg = ctx.enumerate(chrs, k)
    print("enumerate (pageable host buffers): %.3f ms" % ((time.perf_counter() - t0) * 1e3))
