ncu --set full --clock-control none --import-source on -k regex:getrs_dmma -s 0 -c 1 -o gpurun_out/getrs512 -f python tools/run_config.py 512 1000 16 1 x > gpurun_out/ncu5.log 2>&1
tail -n 2 gpurun_out/ncu5.log
