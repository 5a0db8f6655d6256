#!/bin/bash
for n in 512 448; do
  b=$((4000*512*512/n/n))
  TIER=7 timeout 120 python tools/run_config.py $n $b 0 3 | tail -1
  timeout 120 python tools/run_config.py $n $b 0 3 | tail -1
done
