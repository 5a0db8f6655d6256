#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r38_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r38_tests.log
tail -3 gpurun_out/r38_tests.log
for n in 33 40 44; do
  b=$((50000*128*128/n/n))
  timeout 60 python tools/run_config.py $n $b 0 3 | tail -1
done
