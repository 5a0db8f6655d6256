#!/bin/bash
# usage: tools/sass_hist.sh <file.cu> <kernel-mangled-substring> [n]
cd /root/repo/magma_b200/csrc
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xptxas -v -Xcompiler -fPIC -I../../include -c $1 -o /tmp/sh.o 2>&1 | grep -E "error|warning: |Compiling|registers|spill" | paste - - - | sed 's/ptxas info    : //g' | grep -E "$2|error" | cut -c 60-420
cuobjdump -sass /tmp/sh.o | awk -v k="$2" '/Function : /{p=0} $0 ~ "Function : .*"k {p=1} p' > /tmp/sh.sass
grep -cE "^\s+/\*[0-9a-f]{4,5}\*/" /tmp/sh.sass
grep -E "^\s+/\*[0-9a-f]{4,5}\*/" /tmp/sh.sass | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//; s/@!?U?P[0-9T] //' | awk '{print $1}' | sed 's/;//' | sort | uniq -c | sort -rn | head -${3:-30}
