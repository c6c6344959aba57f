"""Dev tool: kernel-time sweep over the hash-partition size (SIBGPU_PART_RECORDS) on one GPU."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sibelia_b200 as sb
from sibelia_b200 import synth

mb = float(sys.argv[1]) if len(sys.argv) > 1 else 100
k = int(sys.argv[2]) if len(sys.argv) > 2 else 25
kind = sys.argv[3] if len(sys.argv) > 3 else "random"
if kind == "random":
    g = [synth.random_genome(int(mb * 1e6), 12345)]
else:
    g = synth.strains(4, int(mb * 1e6 / 4))
for target in [int(x) for x in (sys.argv[4].split(",") if len(sys.argv) > 4 else
                                 "1048576,2097152,4194304,8388608".split(","))]:
    os.environ["SIBGPU_PART_RECORDS"] = str(target)
    c = sb.Context(0)
    c.upload(g)
    c.set_profiling(True)
    for _ in range(3):
        cnt, ninst = c.enumerate_resident(k)
    st = {s["name"]: s for s in c.kernel_stats()}
    tot = c.last_device_ms()
    print("target %9d  total %7.3f ms  V=%d I=%d | " % (target, tot, cnt, ninst) + "  ".join(
        "%s %.3f(%d)" % (n, st[n]["ms"], st[n]["launches"]) for n in
        ("k_scan_hist", "k_scatter", "k_insert", "k_table_scan", "k_mark", "k_emit") if n in st), flush=True)
    c.close()
