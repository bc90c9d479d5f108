import csv, sys
rows=[]
with open(sys.argv[1]) as f:
    lines=[l for l in f if not l.startswith('==')]
for row in csv.DictReader(lines):
    rows.append((row['Kernel Name'], float(row['Metric Value'].replace(',',''))))
names=[x[0] for x in rows]
idx=[i for i,nm in enumerate(names) if 'stepS' in nm]
lo,hi=idx[-2]+1, idx[-1]+1
tot=0
agg={}
for nm,v in rows[lo:hi]:
    key=nm.split('(')[0][-60:]
    agg.setdefault(key,[0,0]); agg[key][0]+=v; agg[key][1]+=1
    tot+=v
for k,(v,c) in sorted(agg.items(), key=lambda kv:-kv[1][0]):
    print(f'{v/1000:9.1f} us x{c:2d}  {k}')
print('iteration total us', tot/1000, 'kernels', hi-lo)
