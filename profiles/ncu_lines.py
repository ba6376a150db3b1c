"""aggregate an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line / region"""
import csv, sys
path=sys.argv[1]; ntiles=float(sys.argv[2]) if len(sys.argv)>2 else 1
rows=list(csv.reader(open(path)))
cur=None; hdr=None; per={}
for r in rows:
    if not r: continue
    if r[0]=="File Path": cur=r[1].split('/')[-1]; continue
    if r[0]=="Function Name": continue
    if r[0]=="Line No": hdr=r; ii=hdr.index("Instructions Executed"); isamp=hdr.index("# Samples"); continue
    if hdr is None: continue
    try: ln=int(r[0])
    except: continue
    if r[2]!="-": continue   # sass rows repeat; keep the per-line summary rows (address "-")
    try: n=int(r[ii]); s=int(r[isamp] or 0)
    except: continue
    per[(cur,ln)]=(n,s,r[1])
tot=sum(v[0] for v in per.values()); stot=sum(v[1] for v in per.values())
print("total inst",tot,"per tile",tot/ntiles,"samples",stot)
top=sorted(per.items(), key=lambda kv:-kv[1][0])[:int(sys.argv[3]) if len(sys.argv)>3 else 60]
for (f,ln),(n,s,src) in top:
    print(f"{f}:{ln:4d} inst {n:>10} {100*n/tot:5.1f}%  /tile {n/ntiles:7.1f}  smp {100*s/max(stot,1):5.1f}%  | {src.strip()[:110]}")
import json
json.dump({f"{f}:{ln}":[n,s] for (f,ln),(n,s,_) in per.items()}, open('/tmp/ncu_lines.json','w'))
