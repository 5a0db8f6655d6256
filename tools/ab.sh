#!/bin/bash
# same-box A/B of two builds: tools/ab.sh "n batch" ...   (A = lib/libmagma_b200_prev.so, B = current)
for cfg in "$@"; do
  for rep in 1 2; do
    a=$(MB200_LIB=$PWD/magma_b200/lib/libmagma_b200_prev.so python tools/fused_time.py $cfg 3 | sort -t' ' -k3 -n | head -1)
    b=$(python tools/fused_time.py $cfg 3 | sort -t' ' -k3 -n | head -1)
    echo "A(prev) $a | B(new) $b"
  done
done
