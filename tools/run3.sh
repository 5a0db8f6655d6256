python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for sr in 0 4 6; do echo "SMALL_ROWS=$sr"; SMALL_ROWS=$sr python tools/run_config.py 16 1000000 1 4 | tail -2; SMALL_ROWS=$sr python tools/run_config.py 16 1000000 0 4 | tail -1; done
python tools/run_config.py 8 2000000 0 3 | tail -1
python tools/run_config.py 32 400000 0 3 | tail -1
SMALL_ROWS=3 python -m pytest tests -m gpu -x -q -k "small or gesv or golden or singular or full_size" 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:lu_sq -s 0 -c 1 -o gpurun_out/sq16_gesv -f python tools/run_config.py 16 200000 1 1 > gpurun_out/ncu1.log 2>&1
