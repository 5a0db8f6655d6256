"""Region view of an ncu source page: executed instructions and stall samples per SASS chunk.
   python tools/ncu_regions.py file.ncu-rep nmatrices [nchunks]"""
import csv, io, subprocess, collections, sys
rep=sys.argv[1]; nm=float(sys.argv[2]); nch=int(sys.argv[3]) if len(sys.argv)>3 else 45
src = subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows=[r for r in csv.reader(io.StringIO(src)) if len(r)>5]
h=rows[0]; ci=h.index("Source"); cx=h.index("Instructions Executed"); cs=h.index("# Samples")
names=[x for x in h if x.startswith("stall_") and "Not Issued" not in x]; idx={x:h.index(x) for x in names}
tot=sum(float(r[cx] or 0) for r in rows[1:]); ts=sum(float(r[cs] or 0) for r in rows[1:])
print("total warp instr per matrix", tot/nm, "samples", ts)
n=len(rows)-1; chunk=max(1,n//nch)
for i in range(1,n,chunk):
    seg=rows[i:i+chunk]
    ex=sum(float(r[cx] or 0) for r in seg); sm=sum(float(r[cs] or 0) for r in seg)
    why=collections.Counter()
    for r in seg:
        for x in names: why[x]+=float(r[idx[x]] or 0)
    w=why.most_common(2)
    print(f"instr {i:5d}: exec/matrix {ex/nm:8.0f}  samples {100*sm/ts:5.1f}%  {w[0][0]}={w[0][1]:.0f} {w[1][0]}={w[1][1]:.0f}  first: {seg[0][ci].strip()[:40]}")
