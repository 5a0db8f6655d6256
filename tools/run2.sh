python -m pytest tests -m gpu -x -q 2>&1 | tail -5
PROBE_SMALL=1 python tools/gpu_probe.py 2>&1 | tee gpurun_out/probe_small.jsonl
ncu --set full --clock-control none --import-source on -k regex:lu_sq -s 0 -c 1 -o gpurun_out/sq16_gesv -f python tools/run_config.py 16 200000 1 1 > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lu_sq -s 0 -c 1 -o gpurun_out/sq32_getrf -f python tools/run_config.py 32 100000 0 1 > gpurun_out/ncu2.log 2>&1
tail -n 3 gpurun_out/ncu1.log gpurun_out/ncu2.log
