#!/bin/bash
# round-2 end-of-round run on the GPU box: full GPU test suite, both bench arms, ncu launch list of the bench command,
# one full capture of the headline kernel. Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r02_tests.log
tail -3 gpurun_out/r02_tests.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r02.json 2> gpurun_out/bench_r02.err
python bench.py > gpurun_out/bench_r02.json 2>> gpurun_out/bench_r02.err; tail -c 300 gpurun_out/bench_r02.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-ref > gpurun_out/bench_under_ncu_r02.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lu_sqs -s 3 -c 1 -o gpurun_out/headline_r02 -f python bench.py --steps 2 --warmup 3 --no-cpu --no-sweep --no-ref > gpurun_out/ncu_headline_r02.log 2>&1
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r02.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["copy_only_ms"], d["cpu_baseline"]["value"])
for r in d.get("sweep",[]): print(r["config"], round(r["ms"],3), round(r["gflops"]), r.get("frac_of_roofline"), r.get("getrs"))
for r in d.get("ref_gpu") or []: print(r)
r=json.loads(open("gpurun_out/bench_ref_r02.json").read().strip().splitlines()[-1]); print("ref arm", r["value"], r["config"]["workload"])
print("ours workload", d["config"]["workload"])
PY
