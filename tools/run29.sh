#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:left_update --launch-skip 2 --launch-count 1 -o gpurun_out/left3_128 -f python tools/run_config.py 128 9472 0 1 > gpurun_out/r29_ncu.log 2>&1
tail -2 gpurun_out/r29_ncu.log
