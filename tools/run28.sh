#!/bin/bash
for n in 128 256 512; do
  b=$((4000*512*512/n/n))
  timeout 60 python tools/run_config.py $n $b 0 3 | tail -1
done
