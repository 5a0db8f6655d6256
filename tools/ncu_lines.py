"""Per-instruction view of an ncu source page: python tools/ncu_lines.py rep start count"""
import csv, io, subprocess, sys
rep=sys.argv[1]; st=int(sys.argv[2]); n=int(sys.argv[3])
src = subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows=[r for r in csv.reader(io.StringIO(src)) if len(r)>5]
h=rows[0]; ci=h.index("Source"); cx=h.index("Instructions Executed"); cs=h.index("# Samples")
names=[x for x in h if x.startswith("stall_") and "Not Issued" not in x]; idx={x:h.index(x) for x in names}
for i,r in enumerate(rows[1+st:1+st+n]):
    why=sorted(((float(r[idx[x]] or 0), x.replace('stall_','')) for x in names), reverse=True)[:2]
    w=" ".join(f"{b}={a:.0f}" for a,b in why if a>0)
    print(f"{st+i:5d} {float(r[cs] or 0):5.0f} x{r[cx]:>6s} {r[ci].strip()[:62]:62s} {w}")
