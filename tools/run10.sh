#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "blocked or rect or getrf" > gpurun_out/r10_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r10_tests.log
tail -5 gpurun_out/r10_tests.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:left_update --launch-skip 7 --launch-count 1 -o gpurun_out/left8 -f python tools/run_config.py 512 592 0 1 > gpurun_out/r10_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r10_launches.csv python tools/run_config.py 512 4000 0 1 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r10_launches.csv 2>/dev/null | head -20
