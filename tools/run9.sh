#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r9_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r9_tests.log
tail -15 gpurun_out/r9_tests.log
for n in 512 256 160; do
  b=$((4000*512*512/n/n))
  timeout 120 python tools/run_config.py $n $b 0 3 | tail -1
  TIER=6 timeout 120 python tools/run_config.py $n $b 0 3 | tail -1
done
