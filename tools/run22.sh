#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "blocked or rect or getrf or left or right or vbatched" > gpurun_out/r22_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r22_tests.log
tail -3 gpurun_out/r22_tests.log
for n in 128 112 96 80 64 48; do
  b=$((50000*128*128/n/n))
  timeout 120 python tools/run_config.py $n $b 0 3 | tail -1
  MID_MAX=32 timeout 120 python tools/run_config.py $n $b 0 3 | tail -1
done
