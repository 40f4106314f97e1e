import csv, collections, re, sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
hdr=None; data=[]
for r in rows:
    if r[0]=='ID': hdr=r; continue
    if hdr and r[0].isdigit(): data.append(dict(zip(hdr,r)))
agg=collections.OrderedDict()
def short(n):
    n=re.sub(r'\(.*','',n).replace('void ppasr::','').replace('ppasr::','')
    return n[:100]
for d in data:
    k=short(d['Kernel Name']); v=float(d['Metric Value'])
    if d['Metric Unit']=='ns': v/=1000.0
    elif d['Metric Unit']=='ms': v*=1000.0
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(a[1] for a in agg.values())
print("launches", len(data), "total us", round(tot,1))
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print(f"{a[1]/tot*100:5.1f}%  n={a[0]:4d}  avg={a[1]/a[0]:8.1f}us  {k}")
