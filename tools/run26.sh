#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r26_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r26_tests.log
tail -3 gpurun_out/r26_tests.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:left_update --launch-skip 7 --launch-count 1 -o gpurun_out/left8d -f python tools/run_config.py 512 592 0 1 > gpurun_out/r26_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_n512_left.csv python tools/run_config.py 512 4000 0 1 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/launches_n512_left.csv | head -12
