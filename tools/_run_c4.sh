timeout 600 python -m pytest tests/test_gpu_round2.py -x -q -k "any_alignment" 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4.csv python tools/vbatched_time.py 1 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/launches_c4.csv | head -30
