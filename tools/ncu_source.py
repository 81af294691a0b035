"""Compact per-instruction view of an `ncu --page source --csv` dump (development aid)."""
import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hi=[i for i,r in enumerate(rows) if r and r[0]=='Address'][0]
hdr=rows[hi]; idx={h:i for i,h in enumerate(hdr)}
data=[r for r in rows[hi+1:] if len(r)>=len(hdr)-5 and r[0].strip() and r[0]!="Address" and r[0]!="Kernel Name"]
f=lambda r,k: float((r[idx[k]] if idx[k]<len(r) else '0') or 0)
tot_inst=sum(f(r,'Instructions Executed') for r in data); tot_samp=sum(f(r,'# Samples') for r in data)
print('total inst',tot_inst,'samples',tot_samp,'rows',len(data))
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
thr=float(sys.argv[2]) if len(sys.argv)>2 else 0.004
for r in data:
    inst=f(r,'Instructions Executed'); samp=f(r,'# Samples')
    if inst/tot_inst<thr and samp/tot_samp<thr*1.5: continue
    st=sorted(((f(r,s),s) for s in stalls),reverse=True)[:2]
    print('%s %5.1f%%i %5.1f%%s thr %5.1f  %-70s %s'%(r[0][-5:],100*inst/tot_inst,100*samp/tot_samp,f(r,'Avg. Threads Executed'),r[idx['Source']][:70],' '.join('%s:%d'%(s[6:],v) for v,s in st if v>0)))
