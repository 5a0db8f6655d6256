python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/run_config.py 512 4000 0 3 | tail -1
TIER=5 python tools/run_config.py 512 4000 0 2 | tail -1
python tools/run_config.py 256 8000 0 2 | tail -1
python tools/run_config.py 384 4000 0 2 | tail -1
