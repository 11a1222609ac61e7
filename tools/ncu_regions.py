import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
nrays=float(sys.argv[2]) if len(sys.argv)>2 else 268282112
hdr=rows[1]; data=rows[2:]
isrc=hdr.index('Source'); iex=hdr.index('Instructions Executed'); ith=hdr.index('Thread Instructions Executed'); ismp=hdr.index('# Samples')
warps=nrays/32
tot=sum(int(r[iex]) for r in data)
print('total warp-inst per 32 rays', tot/warps, 'samples', sum(int(r[ismp]) for r in data))
groups=[]
for n,r in enumerate(data):
    ex=int(r[iex])/warps; th=int(r[ith]); sm=int(r[ismp])
    if groups and abs(groups[-1]['ex']-ex)<=0.12*max(groups[-1]['ex'],0.02):
        g=groups[-1]; g['n']+=1; g['sum']+=ex; g['th']+=th; g['end']=n; g['smp']+=sm
    else:
        groups.append(dict(start=n,end=n,ex=ex,n=1,sum=ex,th=th,smp=sm,first=r[isrc][:60]))
tsm=sum(g['smp'] for g in groups)
for g in groups:
    if g['sum']>4 or g['smp']>tsm*0.01: print(f"{g['start']:5d}-{g['end']:5d} n={g['n']:4d} ex/warp={g['ex']:7.3f} sum={g['sum']:8.1f} thr={g['th']/max(g['sum']*warps,1):5.1f} stall%={100*g['smp']/tsm:5.1f} {g['first']}")
