SMALL_ROWS=2 python -m pytest tests -m gpu -x -q -k "small or golden or singular" 2>&1 | tail -2
for sr in 0 2; do echo "SMALL_ROWS=$sr"; SMALL_ROWS=$sr python tools/run_config.py 32 1000000 0 3 | tail -1; SMALL_ROWS=$sr python tools/run_config.py 32 10000 0 5 | tail -2; done
