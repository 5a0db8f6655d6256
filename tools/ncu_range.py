"""Aggregate stall samples by opcode and by reason over an instruction index range of an ncu source page.
   python tools/ncu_range.py rep start end"""
import csv, io, subprocess, sys, collections
rep=sys.argv[1]; st=int(sys.argv[2]); en=int(sys.argv[3])
src = subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows=[r for r in csv.reader(io.StringIO(src)) if len(r)>5]
h=rows[0]; ci=h.index("Source"); cx=h.index("Instructions Executed"); cs=h.index("# Samples")
names=[x for x in h if x.startswith("stall_") and "Not Issued" not in x]; idx={x:h.index(x) for x in names}
byop=collections.Counter(); byre=collections.Counter(); tot=0; ex=0
for r in rows[1+st:1+en]:
    t=r[ci].strip().split()
    if not t: continue
    op=t[1] if t[0].startswith('@') and len(t)>1 else t[0]
    v=float(r[cs] or 0); byop['.'.join(op.split('.')[:2])]+=v; tot+=v; ex+=float(r[cx] or 0)
    for x in names: byre[x]+=float(r[idx[x]] or 0)
print("samples",tot,"executed",ex)
print("by opcode:", [(o,int(v)) for o,v in byop.most_common(18)])
print("by reason:", [(o.replace('stall_',''),int(v)) for o,v in byre.most_common(10)])
