"""Warp-stall samples of one kernel from an ncu source page, grouped by code segment (segments end at barriers,
atomics, sleeps and exits) plus the most-sampled instructions with their top stall reasons.

    ncu -i prof.ncu-rep --page source --csv --kernel-name regex:<kernel> > src.csv
    python tools/ncu_stalls.py src.csv [top_n]
"""
import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; idx={h:i for i,h in enumerate(hdr)}
data=[r for r in rows[2:] if len(r)==len(hdr) and r[0].startswith("0x")]
tot=sum(int(r[idx["# Samples"]] or 0) for r in data)
print("total samples",tot, "instructions", len(data))
stalls=[h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
# segment by markers
seg=0; acc=0; segs=[]; start=0
cum=0
for n,r in enumerate(data):
    s=int(r[idx["# Samples"]] or 0); cum+=s
    src=r[idx["Source"]].strip()
    key=src.split()[0] if src else ""
    if any(k in src for k in ("BAR.SYNC","NANOSLEEP","ATOMG","EXIT")) or n==len(data)-1:
        segs.append((start,n,cum)); start=n+1
prev=0
for a,b,c in segs:
    print("instr %4d..%4d  samples %6d (%.1f%%)  ends with: %s"%(a,b,c-prev,100.0*(c-prev)/max(tot,1), data[b][idx["Source"]].strip()[:60])); prev=c
print("--- top instructions")
top=sorted(range(len(data)), key=lambda n:-int(data[n][idx["# Samples"]] or 0))[:int(sys.argv[2]) if len(sys.argv)>2 else 25]
for n in sorted(top):
    r=data[n]; s=int(r[idx["# Samples"]] or 0)
    st=sorted(((int(r[idx[h]] or 0),h) for h in stalls), reverse=True)[:3]
    print("%4d %6d %5.1f%%  %-70s %s"%(n,s,100.0*s/tot,r[idx["Source"]].strip()[:70], " ".join("%s=%d"%(h[6:],v) for v,h in st if v)))
