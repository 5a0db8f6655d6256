set -x
python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; tail -c 600 gpurun_out/bench_r01.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r01.json 2>>gpurun_out/bench_r01.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lu_sqs -s 3 -c 1 -o gpurun_out/headline_r01 -f python bench.py --steps 2 --warmup 3 --no-cpu --no-sweep > gpurun_out/ncu_headline.log 2>&1
tail -n 3 gpurun_out/ncu_headline.log
