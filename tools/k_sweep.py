"""Dev tool (GPU box): BASELINE configs[4] -- k sweep on the 500 MB 4-strain set (SURVEY 8(d) recipe), index only.
Uploads once, then times sibgpu_enumerate_resident per k (second run of each k, device time by CUDA events) and prints
the per-kernel split and the algorithmic HBM rate (DESIGN.md section 4 bytes per base)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sibelia_b200 as sb
from sibelia_b200 import synth

n_strains = int(sys.argv[1]) if len(sys.argv) > 1 else 4
base_len = int(float(sys.argv[2])) if len(sys.argv) > 2 else 125_000_000
chrs = synth.strains(n_strains, base_len)
N = sum(len(c) for c in chrs)
ctx = sb.Context(0)
ctx.upload(chrs)
ctx.set_profiling(True)
print("# %d strains x %d bases = %d bases; device-resident text; ms = CUDA events around the whole enumeration" % (n_strains, base_len, N))
print("# %6s %10s %12s %9s %10s %9s  top kernels (ms)" % ("k", "vertices", "instances", "ms", "Gbases/s", "algoGB/s"))
for k in [int(x) for x in os.environ.get("KSWEEP_K", "15,25,100,500,5000").split(",")]:
    for rep in range(2):
        count, ninst = ctx.enumerate_resident(k)
    ms = ctx.last_device_ms()
    st = ctx.kernel_stats()
    algo = sum(s["algo_bytes"] for s in st)
    top = sorted(st, key=lambda s: -s["ms"])[:int(os.environ.get("KSWEEP_TOP", "5"))]
    print("  %6d %10d %12d %9.2f %10.2f %9.0f  %s" % (k, count, ninst, ms, N / ms / 1e6, algo / ms / 1e6,
          ", ".join("%s %.2f" % (s["name"], s["ms"]) for s in top)), flush=True)
