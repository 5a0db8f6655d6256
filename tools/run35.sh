#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:left_update --launch-skip 7 --launch-count 1 -o gpurun_out/left8e -f python tools/run_config.py 512 592 0 1 > gpurun_out/r35_ncu.log 2>&1
tail -1 gpurun_out/r35_ncu.log
