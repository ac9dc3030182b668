"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv) per
kernel: launches, total / average duration, DRAM bytes per launch and the DRAM rate they imply.
    python tools/agg_launches.py launches.csv [top] [first_id last_id]"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4]) if len(sys.argv) > 4 else 1 << 60
per = collections.defaultdict(dict)
names = {}
for x in csv.DictReader(lines):
    i = int(x["ID"])
    if i < lo or i > hi:
        continue
    v = float(x["Metric Value"].replace(",", ""))
    u = x["Metric Unit"]
    if x["Metric Name"].startswith("gpu__time"):
        v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v * 1e6 if u in ("s", "second") else v
        per[i]["us"] = v
    else:
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        per[i]["rd" if "read" in x["Metric Name"] else "wr"] = v * mult
    n = re.sub(r"\(.*", "", x["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
    names[i] = n
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for i, d in per.items():
    a = agg[names[i]]
    a[0] += 1
    a[1] += d.get("us", 0.0)
    a[2] += d.get("rd", 0.0)
    a[3] += d.get("wr", 0.0)
tot = sum(a[1] for a in agg.values())
print("%d launches, %.2f ms in total" % (len(per), tot / 1e3))
print("%9s %6s %5s %9s %9s %9s %8s  kernel" % ("ms", "%", "n", "us/launch", "rd MB", "wr MB", "GB/s"))
for n, (c, t, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%9.3f %5.1f%% %5d %9.1f %9.1f %9.1f %8.0f  %s" % (t / 1e3, 100 * t / tot, c, t / c, rd / c / 1e6, wr / c / 1e6, (rd + wr) / (t * 1e-6) / 1e9 if t else 0, n[:90]))
