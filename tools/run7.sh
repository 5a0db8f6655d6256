#!/bin/bash
# two-phase Tier M: parity + timing A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "mid or vbatched or nonsquare or rect" > gpurun_out/r7_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r7_tests.log
tail -5 gpurun_out/r7_tests.log
for n in 64 96 128; do
  timeout 120 python tools/run_config.py $n 50000 0 3
  SMALL_ROWS=8 timeout 120 python tools/run_config.py $n 50000 0 3
done
