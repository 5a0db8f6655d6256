for rows in 0 5 6 0 5 6; do python tools/small_time.py 32 1000000 $rows; done
for rows in 0 5; do python tools/small_time.py 32 10000 $rows 20; done
echo "--- headline A (prev) / B (new)"
for rep in 1 2; do
MB200_LIB=$PWD/magma_b200/lib/libmagma_b200_prev.so python bench.py --no-sweep --no-cpu --no-ref --steps 20 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('A', d['ms_per_step'])"
python bench.py --no-sweep --no-cpu --no-ref --steps 20 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B', d['ms_per_step'])"
done
python tools/small_time.py 16 1000000 0; MB200_LIB=$PWD/magma_b200/lib/libmagma_b200_prev.so python tools/small_time.py 16 1000000 0
python tools/small_time.py 8 1000000 0; MB200_LIB=$PWD/magma_b200/lib/libmagma_b200_prev.so python tools/small_time.py 8 1000000 0
