"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.

    python tools/launch_summary.py gpurun_out/launches.csv > profiles/rN_launches_summary.txt
"""
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, iv, ig, ib = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
agg, order = {}, []
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("rd3::", "")
    if name not in agg:
        agg[name] = [0, 0.0, ""]
        order.append(name)
    a = agg[name]
    a[0] += 1
    a[1] += float(r[iv].replace(",", "")) / 1e3
    a[2] = "%s x %s" % (r[ig], r[ib])
tot = sum(a[1] for a in agg.values())
print("%-44s %9s %12s %8s  %s" % ("kernel", "launches", "total_us", "share", "grid x block (last seen)"))
for n in order:
    c, t, g = agg[n]
    print("%-44s %9d %12.1f %7.1f%%  %s" % (n[:44], c, t, 100 * t / tot, g))
print("%-44s %9d %12.1f" % ("TOTAL", sum(a[0] for a in agg.values()), tot))
