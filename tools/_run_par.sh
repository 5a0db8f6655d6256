timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -x -q 2>&1 | tail -2
for cfg in "127 50000" "128 50000" "511 2000" "512 2000" "96 50000" "95 50000"; do
for rep in 1 2; do
python tools/fused_time.py $cfg 3 | sort -t' ' -k3 -n | head -1; MB200_LIB=$PWD/magma_b200/lib/libmagma_b200_prev.so python tools/fused_time.py $cfg 3 | sort -t' ' -k3 -n | head -1
done; done
python tools/vbatched_time.py 3 | tail -1
