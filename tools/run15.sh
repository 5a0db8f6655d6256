#!/bin/bash
mkdir -p gpurun_out
TIER=7 timeout 300 ncu --set full --clock-control none --import-source on -k regex:left_update --launch-skip 14 --launch-count 1 -o gpurun_out/left15 -f python tools/run_config.py 512 592 0 1 > gpurun_out/r15_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:left_update --launch-skip 6 --launch-count 1 -o gpurun_out/left7_256 -f python tools/run_config.py 256 2368 0 1 >> gpurun_out/r15_ncu.log 2>&1
tail -3 gpurun_out/r15_ncu.log
