#!/bin/bash
for c in 0 32 48 64 128 256 1000; do
  echo "L2 chunk $c"; MB200_L2_CHUNK=$c timeout 120 python tools/run_config.py 512 4000 0 3 | tail -1
done
for c in 0 128 512; do
  echo "L2 chunk $c n=256"; MB200_L2_CHUNK=$c timeout 120 python tools/run_config.py 256 16000 0 3 | tail -1
done
