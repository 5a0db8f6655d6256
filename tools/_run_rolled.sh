timeout 300 python -m pytest tests/test_gpu_round2.py -x -q -k "fused or chain_panel" 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_n128_b.csv python tools/fused_time.py 128 50000 1 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:panel_chain_kernel -c 1 -o gpurun_out/panel_chain_rolled -f python tools/fused_time.py 128 50000 1 > /dev/null 2>&1
ls -la gpurun_out/
