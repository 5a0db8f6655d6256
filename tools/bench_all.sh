#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r17_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r17_tests.log
tail -3 gpurun_out/r17_tests.log
python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; tail -c 300 gpurun_out/bench_r01.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r01.json 2>>gpurun_out/bench_r01.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_n256_r01.csv python tools/run_config.py 256 16000 0 1 > /dev/null 2>&1
python -c "
import json
d=json.loads(open('gpurun_out/bench_r01.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline'], d['e2e'], d['cpu_baseline'])
for r in d.get('sweep',[]): print(r['config'], r['ms'], r['gflops'], r.get('frac_of_roofline'), r.get('getrs'))
"
