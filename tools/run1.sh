set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
./tools/dmma_probe > gpurun_out/dmma_probe.json; cat gpurun_out/dmma_probe.json
python tools/gpu_probe.py > gpurun_out/probe_r01b.jsonl 2>&1; cat gpurun_out/probe_r01b.jsonl
ncu --set full --clock-control none --import-source on -k regex:lu_small -s 0 -c 1 -o gpurun_out/small16_gesv python tools/run_config.py 16 200000 1 1 > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lu_small -s 0 -c 1 -o gpurun_out/small32_getrf python tools/run_config.py 32 100000 0 1 > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu1.log gpurun_out/ncu2.log
