python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/run_config.py 16 1000000 1 4 | tail -2
python tools/run_config.py 16 1000000 0 3 | tail -1
python tools/run_config.py 8 2000000 0 3 | tail -1
