"""Dev tool: time sibgpu_simplify against the reference stage on synthetic strains (needs oracle/_ref on the box)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sibelia_b200 as sb
from sibelia_b200 import synth
from oracle import ref

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 4
bl = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000
ps = float(sys.argv[3]) if len(sys.argv) > 3 else 0.002
stages = [(25, 150), (100, 1000), (1000, 5000), (5000, 15000)]
chrs = [c.tobytes() for c in synth.strains(ns, bl, p_sub=ps)]
op = [np.arange(len(c), dtype=np.uint32) for c in chrs]
ctx = sb.Context(0)
ctx.simplify(chrs[:1], op[:1], 25, 150, 1)   # warm-up
rchrs, rop = chrs, op
for (k, D) in stages:
    t0 = time.perf_counter()
    g = ctx.simplify(chrs, op, k, D, 4)
    tg = time.perf_counter() - t0
    dev_ms = ctx.last_device_ms()
    line = "stage (%d,%d): ours %.3f s (device %.1f ms, %d launches), bulges %d" % (k, D, tg, dev_ms, ctx.last_launches(), g[2])
    if ref.available() and "--noref" not in sys.argv:
        r = ref.simplify(rchrs, rop, k, D, 4)
        ok = r[2] == g[2] and all(a == b for a, b in zip(r[0], g[0])) and all(np.array_equal(a, b) for a, b in zip(r[1], g[1]))
        line += " | reference %.3f s, bulges %d | identical=%s | speed-up %.1fx" % (r[3], r[2], ok, r[3] / tg)
        rchrs, rop = r[0], r[1]
    print(line, flush=True)
    chrs, op = g[0], g[1]
