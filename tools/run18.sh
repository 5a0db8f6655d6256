#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "blocked or rect or getrf or left or right or vbatched" > gpurun_out/r18_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r18_tests.log
tail -3 gpurun_out/r18_tests.log
for n in 512 384 256 160; do
  b=$((4000*512*512/n/n))
  timeout 120 python tools/run_config.py $n $b 0 3 | tail -1
done
