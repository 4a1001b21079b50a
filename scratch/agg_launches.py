import csv, collections, re, sys
lines=[l for l in open(sys.argv[1]) if not l.startswith('==')]
r=csv.DictReader(lines)
agg=collections.OrderedDict()
for row in r:
    name=re.sub(r'\(.*','',row['Kernel Name'])[:64]
    v=float(row['Metric Value'].replace(',',''))
    u=row['Metric Unit']
    v = v/1e6 if u=='ns' else v/1e3 if u=='us' else v*1e3 if u in('second','s') else v
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(v[1] for v in agg.values())
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:int(sys.argv[2]) if len(sys.argv)>2 else 18]:
    print("%-66s n=%3d total %9.3f ms avg %8.3f ms %5.1f%%"%(k,n,t,t/n,100*t/tot))
