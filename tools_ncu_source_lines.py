"""Aggregates `ncu -i X.ncu-rep --page source --csv` by source line: share of samples / executed instructions and top stall\nreasons.  usage: ncu -i rep --page source --csv > src.csv; python tools_ncu_source_lines.py src.csv [min_share]"""
import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
thr=float(sys.argv[2]) if len(sys.argv)>2 else 0.006
cur=None; hdr=None
agg={}
tot=0; toti=0
def I(x):
    try: return int(x)
    except: return 0
for r in rows:
    if not r: continue
    if r[0]=="File Path": cur=r[1].split('/')[-1]; continue
    if r[0]=="Line No": 
        hdr=r; si=hdr.index("# Samples"); ii=hdr.index("Instructions Executed")
        stall_cols=[(j,h) for j,h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if r[0]=="Function Name": continue
    if hdr and len(r)>ii and r[0].isdigit():
        line=int(r[0]); s=I(r[si]); i=I(r[ii])
        st={h:I(r[j]) for j,h in stall_cols if I(r[j])}
        agg[(cur,line)]=(s,i,r[1][:100],st)
        tot+=s; toti+=i
print("total samples",tot,"instr",toti)
for (f,l),(s,i,src,st) in sorted(agg.items()):
    if s>tot*thr:
        top=sorted(st.items(), key=lambda x:-x[1])[:3]
        print(f"{f[:16]:16s}{l:4d} {100*s/tot:5.1f}%smp {100*i/toti:5.1f}%ins {src.strip()[:64]:64s} {[(k[6:],v) for k,v in top]}")
