"""Summarise an ncu launch-list csv: time by kernel name (us) + optional per-launch listing.
   python tools/launch_summary.py launches.csv [-v]"""
import csv, collections, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
data = collections.defaultdict(dict); names = {}
for row in csv.DictReader(lines):
    k = int(row["ID"]); names[k] = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("mb200::<unnamed>::", "")
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    if row["Metric Name"] == "gpu__time_duration.sum": v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
    if u == "Gbyte": v *= 1e9
    if u == "Mbyte": v *= 1e6
    if u == "Kbyte": v *= 1e3
    data[k][row["Metric Name"]] = v
tot = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for k in sorted(data):
    d = data[k]; t = tot[names[k]]
    t[0] += 1; t[1] += d.get("gpu__time_duration.sum", 0); t[2] += d.get("dram__bytes_read.sum", 0); t[3] += d.get("dram__bytes_write.sum", 0)
    if "-v" in sys.argv: print(k, names[k][:40], f"{d.get('gpu__time_duration.sum',0):.1f}us")
all_t = sum(t[1] for t in tot.values())
for nme, t in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{t[1]:10.1f} us {100*t[1]/all_t:5.1f}%  x{t[0]:<4d} rd {t[2]/1e9:7.3f} GB wr {t[3]/1e9:7.3f} GB  {nme[:60]}")
