#!/usr/bin/env python
"""Aggregate the SASS-level sampling page of an ncu report per region of consecutive instructions:
    ncu -i rep.ncu-rep --page source --csv > src.csv;  python scripts/ncu_regions.py src.csv [instructions per region]
Per region: share of stall samples, share of executed warp instructions, active lanes, top opcodes, top stall reasons.
This is how the Drucker-Prager tangent-store loop (57 % of the instructions) was found."""
import csv,sys
f=sys.argv[1]; chunk=int(sys.argv[2]) if len(sys.argv)>2 else 60
rows=list(csv.reader(open(f)))
h=rows[1]; data=rows[2:]
iS=h.index('# Samples'); iI=h.index('Instructions Executed'); iT=h.index('Thread Instructions Executed'); isrc=h.index('Source')
stalls=[c for c in h if c.startswith('stall_') and 'Not Issued' not in c]
tot_s=sum(int(r[iS]) for r in data); tot_i=sum(int(r[iI]) for r in data)
print(rows[0][1][:100]); print('rows',len(data),'samples',tot_s,'warp inst',tot_i)
for k in range(0,len(data),chunk):
    seg=data[k:k+chunk]
    s=sum(int(r[iS]) for r in seg); i=sum(int(r[iI]) for r in seg); t=sum(int(r[iT]) for r in seg)
    if i==0 and s==0: continue
    ops={}
    for r in seg:
        toks=r[isrc].split()
        op=toks[1] if toks[0].startswith('@') else toks[0]
        op=op.split('.')[0]; ops[op]=ops.get(op,0)+1
    top=sorted(ops.items(),key=lambda x:-x[1])[:5]
    st={c:sum(int(r[h.index(c)]) for r in seg) for c in stalls}
    tops=sorted(st.items(),key=lambda x:-x[1])[:3]
    print(f'{k:5d} smp {100*s/tot_s:5.1f}% inst {100*i/tot_i:5.1f}% lanes {t/max(i,1):4.1f}', top, [(a[6:],b) for a,b in tops])
