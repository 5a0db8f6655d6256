for rows in 0 2 5 6 0 2; do python tools/small_time.py 32 1000000 $rows; done
for rows in 0 2; do python tools/small_time.py 24 1000000 $rows; python tools/small_time.py 20 1000000 $rows; done
