import csv, collections, sys
path=sys.argv[1]; nsteps=int(sys.argv[2]) if len(sys.argv)>2 else 3
rows=[r for r in csv.reader(open(path)) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size')
data=[(r[ki].split('(')[0], float(r[vi].replace(',','')), r[gi]) for r in rows[1:]]
n=len(data)//nsteps; last=data[-n:]
tot=sum(v for _,v,_ in last)
print('launches/step', n, 'sum us', tot/1000)
for k,v,g in last: print(f'{k[:40]:40s} {v/1000:8.1f} us  grid {g}')
