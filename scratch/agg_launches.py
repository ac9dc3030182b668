import csv, collections, re, sys
rows=[]
with open(sys.argv[1]) as f:
    lines=[l for l in f if l.startswith('"')]
for x in csv.DictReader(lines):
    rows.append(x)
agg=collections.defaultdict(lambda:[0,0.0])
for x in rows:
    n=x['Kernel Name']; n=re.sub(r'\(.*','',n); n=n.replace('void ','').replace('<unnamed>::','')
    v=float(x['Metric Value'].replace(',',''))
    u=x['Metric Unit']
    v = v/1e3 if u=='ns' else v*1e3 if u=='ms' else v*1e6 if u=='s' else v
    agg[n][0]+=1; agg[n][1]+=v
tot=sum(v[1] for v in agg.values())
top=int(sys.argv[2]) if len(sys.argv)>2 else 30
print(f"{len(rows)} launches, total {tot/1e3:.2f} ms (all steps in the capture)")
for n,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:top]:
    print(f"{t/1e3:9.2f} ms {100*t/tot:5.1f}% {c:5d} x {t/c:9.1f} us  {n[:100]}")
