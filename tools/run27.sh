#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "blocked or rect or getrf or left or right or vbatched" > gpurun_out/r27_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r27_tests.log
tail -3 gpurun_out/r27_tests.log
for n in 64 128 256 384 512; do
  b=$((4000*512*512/n/n))
  timeout 60 python tools/run_config.py $n $b 0 3 | tail -1
done
