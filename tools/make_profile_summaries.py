"""Dev tool (runs here, no GPU): turns gpurun_out/<round>_*.ncu-rep + <round>_launches.csv into the tracked summaries
under profiles/ and profiles/traffic.json (per-launch dram bytes of every profiled kernel, read by bench.py).
usage: python tools/make_profile_summaries.py [r1|r2]   (r1: tools/profile_all.sh, warm caches; r2: tools/profile_r2.sh,
--cache-control all)"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RND = sys.argv[1] if len(sys.argv) > 1 else "r1"
CACHE = "none" if RND == "r1" else "all"
OUT = os.path.join(ROOT, "profiles")
GO = os.path.join(ROOT, "gpurun_out")
PATS = [r"^gpu__time_duration\.sum$", r"^dram__bytes_(read|write)\.sum$", r"^lts__t_sector_hit_rate\.pct$",
        r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$", r"^launch__registers_per_thread$", r"^launch__grid_size$",
        r"^launch__block_size$", r"^lts__throughput\.avg\.pct_of_peak_sustained_elapsed$",
        r"^l1tex__throughput\.avg\.pct_of_peak_sustained_elapsed$", r"^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$",
        r"^gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$", r"^dram__throughput\.avg\.pct_of_peak_sustained_elapsed$",
        r"^lts__t_sectors_srcunit_tex_op_(atom|red|read|write)\.sum$", r"^smsp__inst_executed\.sum$",
        r"^smsp__thread_inst_executed_per_inst_executed\.ratio$", r"^sm__inst_executed_pipe_lsu\.avg\.pct_of_peak_sustained_active$",
        r"^l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom\.sum$", r"^smsp__warp_issue_stalled_.*_per_warp_active\.pct$",
        r"^launch__occupancy_limit_(registers|shared_mem|warps)$", r"^launch__shared_mem_per_block_(dynamic|static)$"]
traffic = {}
for f in sorted(os.listdir(GO)):
    if not re.match(RND + r"_\w+\.ncu-rep$", f):
        continue
    raw = subprocess.run(["ncu", "-i", os.path.join(GO, f), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        continue
    hdr, units, alldata = rows[0], rows[1], rows[2:]
    kn = hdr.index("Kernel Name")
    groups = collections.OrderedDict()
    for r in alldata:
        base = re.sub(r"^void\s+", "", r[kn]).split("<")[0].split("(")[0].split("::")[-1]
        groups.setdefault(base, []).append(r)
    for name, data in groups.items():
        if RND != "r1":
            name = name + {"r2_fp.ncu-rep": "_k100"}.get(f, "")
        lines = ["# ncu --set full --clock-control none --cache-control %s, %d launch(es) of %s (report %s)" % (CACHE, len(data), name, f),
                 "# command: see tools/profile_%s.sh (python bench.py --steps 1 --warmup 3: 100 MB random contig, k=25, unless noted)"
                 % ("all" if RND == "r1" else RND), ""]
        lines.append("kernel: " + " | ".join(sorted(set(r[kn] for r in data))))
        vals = {}
        for i, h in enumerate(hdr):
            if any(re.search(p, h) for p in PATS):
                vals[h] = [r[i] for r in data]
                lines.append("%-80s %-10s %s" % (h, units[i], "  ".join(r[i] for r in data)))
        try:
            SC = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
            rd = [float(x.replace(",", "")) * SC.get(units[hdr.index("dram__bytes_read.sum")], 1.0) for x in vals["dram__bytes_read.sum"]]
            wr = [float(x.replace(",", "")) * SC.get(units[hdr.index("dram__bytes_write.sum")], 1.0) for x in vals["dram__bytes_write.sum"]]
            traffic[name] = (sum(rd) + sum(wr)) / len(rd)
            lines.append("")
            lines.append("dram traffic per launch (read+write): %.1f MB" % (traffic[name] / 1e6))
        except Exception as e:
            lines.append("traffic: n/a (%s)" % e)
        open(os.path.join(OUT, "%s_ncu_%s.txt" % (RND, name)), "w").write("\n".join(lines) + "\n")
        print(name, "->", "%s_ncu_%s.txt" % (RND, name), "%.1f MB/launch" % (traffic.get(name, 0) / 1e6))
# bench.py looks kernels up by their profiler-span names
alias = {"k_insert_compact": "k_insert", "k_table_scan_compact": "k_table_scan"}
tj = {alias.get(k, k): v for k, v in traffic.items()}
json.dump(tj, open(os.path.join(OUT, "traffic.json"), "w"), indent=1, sort_keys=True)

# launch list -> per-kernel shares
src = os.path.join(GO, RND + "_launches.csv")
if os.path.exists(src):
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        n = r[kn].split("(")[0]
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(OUT, RND + "_launches_summary.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 900 python bench.py --steps 1 --warmup 3\n")
        f.write("# (4 enumerations of the 100 MB random contig, k=25, + torch fill kernels; cold-cache serialised times: compare SHARES)\n")
        for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-70s launches %4d  total %10.1f us  share %5.1f%%\n" % (n[:70], c, t / 1000, 100 * t / tot))
    import shutil
    shutil.copy(src, os.path.join(OUT, RND + "_launches_bench_100mb_k25.csv"))
    print(open(os.path.join(OUT, RND + "_launches_summary.txt")).read())
