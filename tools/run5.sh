ncu --set full --clock-control none --import-source on -k regex:lu_mid -s 0 -c 1 -o gpurun_out/mid128 -f python tools/run_config.py 128 4000 0 1 > gpurun_out/ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lu_mid -s 0 -c 1 -o gpurun_out/mid64 -f python tools/run_config.py 64 16000 0 1 > gpurun_out/ncu4.log 2>&1
tail -n 2 gpurun_out/ncu3.log gpurun_out/ncu4.log
