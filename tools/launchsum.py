import csv,re,collections,sys
rows=list(csv.reader(open(sys.argv[1])))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; kn=h.index('Kernel Name'); mv=h.index('Metric Value'); mu=h.index('Metric Unit')
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[hi+1:]:
    if len(r)<=mv: continue
    name=re.sub(r"\(.*","",r[kn])
    v=float(r[mv].replace(',',''))
    u=r[mu]
    if u=='ns': v/=1e6
    elif u=='us': v/=1e3
    agg[name][0]+=1; agg[name][1]+=v
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print(f"{v[1]:9.3f} ms {100*v[1]/tot:5.1f}%  n={v[0]:3d}  avg {v[1]/v[0]:7.3f}  {k}")
print('total',tot)
