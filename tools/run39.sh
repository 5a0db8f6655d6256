#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r39_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r39_tests.log
tail -3 gpurun_out/r39_tests.log
for n in 96 160 384 512; do
  b=$((4000*512*512/n/n))
  timeout 60 python tools/run_config.py $n $b 0 3 | tail -1
done
timeout 600 python bench.py --no-cpu --steps 5 --warmup 3 > gpurun_out/r39_bench.json 2> gpurun_out/r39_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r39_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'])
for r in d.get('sweep',[]): print(r['config'], r['ms'], r['gflops'], r.get('frac_of_roofline'))
PY
