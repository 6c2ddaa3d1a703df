import csv,collections,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[2]
si=hdr.index('Warp Stall Sampling (All Samples)')
names=[n for n in hdr if n.startswith('stall_') and 'Not Issued' not in n]
cols={n:hdr.index(n) for n in names}
agg=collections.OrderedDict(); tot=0; files=None; cur=None
for r in rows[3:]:
    if len(r)<60:
        if r and r[0]=='File Path': files=r[1].split('/')[-1]
        continue
    if r[2]=='-':
        cur=(files,int(r[0]),r[1].strip())
        try: v=int(r[si])
        except: v=0
        a=agg.setdefault(cur,[0,collections.Counter()]); a[0]+=v
        for n,i in cols.items():
            try: a[1][n]+=int(r[i])
            except: pass
        tot+=v
print('total samples',tot)
tc=collections.Counter()
for k,(v,c) in agg.items(): tc.update(c)
print(tc.most_common(8))
N=int(sys.argv[2]) if len(sys.argv)>2 else 45
for k,(v,c) in sorted(agg.items(), key=lambda x:-x[1][0])[:N]:
    print(f"{v:6d} {100*v/tot:5.1f}% {k[0]}:{k[1]:>4} {k[2][:80]}  {dict(c.most_common(2))}")
