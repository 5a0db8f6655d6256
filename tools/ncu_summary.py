"""Summarise an .ncu-rep: key raw metrics + top stall reasons + hottest source lines.
   python tools/ncu_summary.py file.ncu-rep [nlines]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; nl = int(sys.argv[2]) if len(sys.argv) > 2 else 12
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
        "sm__inst_executed_pipe_lsu.sum", "smsp__inst_executed_pipe_lsu.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        ]
for k in keys:
    if k in d: print(f"{k:75s} {d[k]:>20s} {u[k]}")
st = [(k, float(v.replace(',', ''))) for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio") or (k.startswith("smsp__average_warp_latency_issue_stalled") and v)]
st = [(k, v) for k, v in st if v == v]
st.sort(key=lambda t: -t[1])
print("-- stall reasons (warps stalled per issue) --")
for k, v in st[:8]: print(f"   {k.replace('smsp__average_warps_issue_stalled_','').replace('smsp__average_warp_latency_issue_stalled_','lat_'):60s} {v:8.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
rows = [r for r in rows if len(r) > 5]
if rows:
    h = rows[0]
    def col(name):
        for i, x in enumerate(h):
            if x.strip() == name: return i
        return None
    ci = col("Source"); cs = col("Warp Stall Sampling (All Samples)") or col("# Samples"); cx = col("Instructions Executed")
    ca = col("Address")
    if cs is not None:
        tot = sum(float(r[cs] or 0) for r in rows[1:] if len(r) > cs)
        top = sorted(rows[1:], key=lambda r: -float(r[cs] or 0))[:nl]
        print(f"-- hottest SASS by stall samples (total {tot:.0f}) --")
        names = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
        idx = {x: h.index(x) for x in names}
        for r in top:
            why = sorted(((float(r[idx[x]] or 0), x) for x in names), reverse=True)[:2]
            print(f"   {float(r[cs] or 0)/max(tot,1)*100:5.1f}%  x{r[cx] if cx is not None else ''}  {r[ci].strip()[:70]:70s} {why[0][1]}={why[0][0]:.0f} {why[1][1]}={why[1][0]:.0f}")
        # aggregate by opcode
        agg = collections.defaultdict(float)
        for r in rows[1:]:
            op = r[ci].strip().split()
            if not op: continue
            o = op[1] if op[0].startswith('@') and len(op) > 1 else op[0]
            agg[o.rstrip(';')] += float(r[cs] or 0)
        print("-- stall samples by opcode --")
        for o, v in sorted(agg.items(), key=lambda t: -t[1])[:14]: print(f"   {v/max(tot,1)*100:5.1f}%  {o}")
