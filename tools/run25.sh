#!/bin/bash
mkdir -p gpurun_out
for a in 3 2 1; do
for n in 128 256 512; do
  b=$((4000*512*512/n/n))
  echo "ahead $a"; MB200_LL_AHEAD=$a timeout 60 python tools/run_config.py $n $b 0 3 | tail -1
done; done
