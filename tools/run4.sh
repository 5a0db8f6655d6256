python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/run_config.py 512 4000 16 3 x | tail -1
python tools/run_config.py 256 8000 16 2 x | tail -1
python tools/run_config.py 128 20000 4 2 x | tail -1
