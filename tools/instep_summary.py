#!/usr/bin/env python
"""Per-kernel sums of an ncu --csv launch list (tools/instep_prof.py): launches, time, DRAM read / written.
usage: python tools/instep_summary.py launches.csv [rows_per_launch]"""
import csv, sys, re, collections
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
iK, iM, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
iID = hdr.index("ID")
per = collections.OrderedDict()
launch = {}
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[iK]).replace("void ", "").replace("<unnamed>::", "")
    v = float(r[iV].replace(",", ""))
    u = r[iU]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
    d = launch.setdefault(r[iID], {"name": name})
    d[r[iM]] = v * scale
for d in launch.values():
    a = per.setdefault(d["name"], [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += d.get("gpu__time_duration.sum", 0.0); a[2] += d.get("dram__bytes_read.sum", 0.0); a[3] += d.get("dram__bytes_write.sum", 0.0)
tot = [sum(a[i] for a in per.values()) for i in range(4)]
print("%-28s %6s %10s %12s %12s %10s" % ("kernel", "n", "us", "MB read", "MB written", "GB/s"))
for k, a in sorted(per.items(), key=lambda kv: -kv[1][1]):
    print("%-28s %6d %10.1f %12.1f %12.1f %10.0f" % (k[:28], a[0], a[1], a[2], a[3], (a[2] + a[3]) / a[1] * 1e3 if a[1] else 0))
print("%-28s %6d %10.1f %12.1f %12.1f %10.0f" % ("total", tot[0], tot[1], tot[2], tot[3], (tot[2] + tot[3]) / tot[1] * 1e3))
