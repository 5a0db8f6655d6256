ncu --set full --clock-control none --import-source on -k regex:swap_trsm -s 2 -c 1 -o gpurun_out/swaptrsm -f python tools/run_config.py 512 4000 0 1 > gpurun_out/ncu6.log 2>&1
tail -n 2 gpurun_out/ncu6.log
