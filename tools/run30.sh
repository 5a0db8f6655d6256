#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_2gpu_r01.json 2> gpurun_out/bench_2gpu_r01.err
tail -c 400 gpurun_out/bench_2gpu_r01.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_2gpu_r01.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e'])
for r in d.get('sweep',[]): print(r['config'], r['ms'], r['gflops'])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 | tail -1 | head -c 400
