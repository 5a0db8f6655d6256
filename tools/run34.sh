#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r34_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r34_tests.log
tail -4 gpurun_out/r34_tests.log
for n in 64 128 256 512; do
  b=$((4000*512*512/n/n))
  timeout 60 python tools/run_config.py $n $b 0 3 | tail -1
done
timeout 60 python tools/run_config.py 1024 500 0 2 | tail -1
