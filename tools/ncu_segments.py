"""Executed instructions and stall samples between consecutive barriers of the kernel last summarised by
tools/ncu_summary.py (phases of a tile loop).  Usage: python tools/ncu_segments.py [first_line last_line]"""
import csv,collections,sys
rows=list(csv.reader(open('/tmp/sass/last_src.csv')))
hdr=rows[1]; data=rows[2:]
iA=hdr.index('Source'); iE=hdr.index('Instructions Executed'); iS=hdr.index('# Samples')
seg=0; acc=collections.Counter(); sam=collections.Counter(); ops=collections.defaultdict(collections.Counter); first={}
for n,r in enumerate(data):
    try: e=int(r[iE]); s=int(r[iS])
    except: continue
    src=r[iA].strip()
    t=src.split(); op=t[1] if t[0].startswith('@') else t[0]
    acc[seg]+=e; sam[seg]+=s; ops[seg][op.split('.')[0]]+=e
    first.setdefault(seg,n)
    if op.startswith('BAR') or op.startswith('SYNCS') : seg+=1
tot=sum(acc.values()); ts=sum(sam.values())
for k in sorted(acc):
    if acc[k]>200000: print(k, first[k], f'{acc[k]/48016:8.1f}/warp-iter {100*acc[k]/tot:5.1f}% samples {100*sam[k]/ts:5.1f}%', dict(ops[k].most_common(9)))
if len(sys.argv)>2:
    for n in range(int(sys.argv[1]),int(sys.argv[2])):
        r=data[n]; print(n, r[iE].rjust(8), r[iS].rjust(5), r[iA][:100])
