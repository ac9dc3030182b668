"""Top stall sites of one kernel from `ncu -i rep --page source --csv` output (SASS view): sample totals per stall
reason and the N instructions with the most samples.   python tools/ncu_source_top.py file.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0      # n-th kernel of the file
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
print(rows[starts[which]][1][:100])
rows = rows[starts[which]:starts[which + 1]]
hdr = rows[1]
body = [r for r in rows[2:] if len(r) == len(hdr) and r[0] != "Address"]
ix = {h: i for i, h in enumerate(hdr)}
samp = ix["# Samples"]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[samp] or 0) for r in body)
print("total samples", tot, " instructions", len(body), " warp-instrs executed", sum(int(r[ix["Instructions Executed"]] or 0) for r in body))
agg = {s: sum(int(r[ix[s]] or 0) for r in body) for s in stalls}
print("by reason:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v * 200 > tot))
order = sorted(range(len(body)), key=lambda i: -int(body[i][samp] or 0))[:n]
for i in sorted(order):
    r = body[i]
    top = sorted(((int(r[ix[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
    print("%5d %5.1f%%  %-70s %s" % (i, 100.0 * int(r[samp] or 0) / max(tot, 1), r[ix["Source"]].strip()[:70], " ".join("%s:%d" % (b, a) for a, b in top if a)))
