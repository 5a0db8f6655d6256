"""Executed warp-instructions by opcode from an .ncu-rep source page.  python tools/ncu_ophist.py rep nmatrices"""
import csv, io, subprocess, collections, sys
rep=sys.argv[1]; nm=float(sys.argv[2])
src = subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows=[r for r in csv.reader(io.StringIO(src)) if len(r)>5]
h=rows[0]; ci=h.index("Source"); cx=h.index("Instructions Executed")
agg=collections.Counter(); tot=0
for r in rows[1:]:
    op=r[ci].strip().split()
    if not op: continue
    o=op[1] if op[0].startswith('@') and len(op)>1 else op[0]
    o=o.rstrip(';'); o='.'.join(o.split('.')[:2])
    v=float(r[cx] or 0); agg[o]+=v; tot+=v
print("total/matrix", tot/nm)
for o,v in agg.most_common(40): print(f"{v/nm:8.1f}  {100*v/tot:5.1f}%  {o}")
