timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_n512_c.csv python tools/fused_time.py 512 4000 1 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/launches_n512_c.csv | head -16
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/launches_n512_c.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
H=rows[hdr]; ix={h:i for i,h in enumerate(H)}
for r in rows[hdr+1:]:
    if len(r)<len(H): continue
    if r[ix['Metric Name']]=='gpu__time_duration.sum' and 'panel' in r[ix['Kernel Name']]:
        print(r[ix['ID']], r[ix['Kernel Name']][:50], r[ix['Block Size']] if 'Block Size' in ix else '', r[ix['Metric Value']])
PY
