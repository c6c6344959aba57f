"""Dev tool: print selected metrics from an `ncu --page raw --csv` export (reads the CSV path given)."""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
pats = sys.argv[2:] or [r"gpu__time_duration\.sum", r"dram__bytes_(read|write)\.sum$", r"lts__t_sector_hit_rate\.pct",
                        r"sm__warps_active\.avg\.pct_of_peak", r"launch__registers_per_thread", r"launch__grid_size",
                        r"dram__throughput\.avg\.pct_of_peak_sustained_elapsed$", r"lts__throughput\.avg\.pct",
                        r"l1tex__throughput\.avg\.pct", r"sm__throughput\.avg\.pct", r"lts__t_sectors_op_(atom|red|read|write)\.sum$",
                        r"issue_stalled.*_per_warp_active\.pct", r"sm__inst_executed\.sum$", r"launch__occupancy_limit",
                        r"lts__t_sectors\.sum$", r"lts__t_sectors_srcunit_tex_op_.*\.sum$"]
for i, h in enumerate(hdr):
    if any(re.search(p, h) for p in pats):
        print("%-90s %s" % (h[-90:], [r[i] for r in rows[2:]]))
