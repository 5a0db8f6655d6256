timeout 300 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -x -q -k "getri" 2>&1 | tail -5
python tools/getri_time.py 64 100000
python tools/getri_time.py 32 400000
python tools/getri_time.py 16 1000000
python tools/getri_time.py 48 100000
