#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "blocked or rect or getrf" > gpurun_out/r12_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r12_tests.log
tail -3 gpurun_out/r12_tests.log
for n in 512 384 256 160; do
  b=$((4000*512*512/n/n))
  timeout 120 python tools/run_config.py $n $b 0 3 | tail -1
done
for n in 512 256; do
b=$((4000*512*512/n/n))
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r12_launches_$n.csv python tools/run_config.py $n $b 0 1 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r12_launches_$n.csv 2>/dev/null | head -4
grep left_update gpurun_out/r12_launches_$n.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' '; echo
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:left_update --launch-skip 7 --launch-count 1 -o gpurun_out/left8c -f python tools/run_config.py 512 592 0 1 > gpurun_out/r12_ncu.log 2>&1
