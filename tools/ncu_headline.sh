#!/bin/bash
# one full ncu capture of the headline kernel inside a bench launch + the launch list of the bench command
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:lu_sqs -s 3 -c 1 -o gpurun_out/headline_r01 -f python bench.py --steps 2 --warmup 3 --no-cpu --no-sweep > gpurun_out/ncu_headline.log 2>&1
tail -n 2 gpurun_out/ncu_headline.log
