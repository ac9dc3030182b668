"""Opcode histogram of one kernel from an `ncu --page source --csv` export: executed warp instructions and stall samples
per SASS opcode.   ncu -i X.ncu-rep --page source --csv --kernel-name regex:K --launch-count 1 > src.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
if len(sys.argv) > 3:                                   # n-th kernel of a multi-kernel export
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
    k = int(sys.argv[3])
    print(rows[starts[k]][1][:90])
    rows = rows[starts[k]:starts[k + 1]]
hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
i_src, i_ex, i_samp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ops, samp = collections.Counter(), collections.Counter()
tot = ts = 0
for r in rows:
    if len(r) <= i_ex or r is hdr or not r[i_ex].isdigit():
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[i_src])
    if not m:
        continue
    op = ".".join(m.group(2).split(".")[:2]) if m.group(2).startswith(("MUFU", "F2F", "F2FP", "LDS", "STS", "LDG", "STG")) else m.group(2).split(".")[0]
    n, s = int(r[i_ex]), int(r[i_samp] or 0)
    ops[op] += n
    samp[op] += s
    tot += n
    ts += s
print("total warp instructions %d, samples %d" % (tot, ts))
for op, n in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 28):
    print("%-14s %11d %5.1f%%   samples %5.1f%%" % (op, n, 100.0 * n / tot, 100.0 * samp[op] / max(ts, 1)))
