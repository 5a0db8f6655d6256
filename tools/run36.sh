#!/bin/bash
for n in 44 40 36 33; do
  b=$((50000*128*128/n/n))
  timeout 60 python tools/run_config.py $n $b 0 3 | tail -1
  MID_MAX=32 timeout 60 python tools/run_config.py $n $b 0 3 | tail -1
done
