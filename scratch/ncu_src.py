import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hi=[i for i,r in enumerate(rows) if len(r)>1 and r[0]=='Address']
start=hi[0]; end=hi[1]-1 if len(hi)>1 else len(rows)
h=rows[start]
ia=h.index('Source'); ie=h.index('Instructions Executed'); isamp=h.index('# Samples'); ith=h.index('Avg. Threads Executed')
body=[r for r in rows[start+1:end] if len(r)>ie and r[ie].isdigit()]
tot=sum(int(r[ie]) for r in body); ts=sum(int(r[isamp]) for r in body)
print("total warp instr",tot,"samples",ts, "nlines", len(body))
n=int(sys.argv[2]) if len(sys.argv)>2 else 40
mode=sys.argv[3] if len(sys.argv)>3 else 'exec'
key=(lambda r:-int(r[ie])) if mode=='exec' else (lambda r:-int(r[isamp]))
for r in sorted(body,key=key)[:n]:
    print("%6.2f%% ex %6.2f%% smp thr %5s  %s"%(100*int(r[ie])/tot,100*int(r[isamp])/max(ts,1),r[ith],r[ia].strip()[:90]))
# opcode histogram weighted by executed
import collections,re
hist=collections.Counter()
for r in body:
    op=r[ia].strip().split()
    if not op: continue
    o=op[0] if not op[0].startswith('@') else (op[1] if len(op)>1 else op[0])
    hist[o.split('.')[0]]+=int(r[ie])
print({k:round(100*v/tot,1) for k,v in hist.most_common(18)})
