"""Dev tool: per-SASS-instruction summary of an `ncu --page source --csv` export: share of samples / instructions,
shared-memory wavefronts.  usage: ncu -i rep --page source --csv --kernel-name regex:K > f.csv; python tools/ncu_source.py f.csv [min_pct]"""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
hdr = None
body = []
for r in rows:
    if r[0] == "Address":
        if hdr is not None:
            break
        hdr = r
    elif hdr is not None:
        body.append(r)
ia, isamp, iinst, iw = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("L1 Wavefronts Shared")
tot = sum(int(r[isamp]) for r in body) or 1
toti = sum(int(r[iinst]) for r in body) or 1
print("total samples", tot, "warp instructions", toti, "SASS lines", len(body))
for n, r in enumerate(body):
    ps, pi = 100 * int(r[isamp]) / tot, 100 * int(r[iinst]) / toti
    if ps >= thr or pi >= 2 * thr:
        print("%4d %-66s samp %5.1f%% inst %5.2f%% wf %s" % (n, r[ia].strip()[:66], ps, pi, r[iw]))
