timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "radix or scan" 2>&1 | tail -3
timeout 300 python scripts/bench_sort.py 1.23e8 2>&1 | tail -14
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/sort_launches.csv python scripts/bench_sort.py 1.23e8 > gpurun_out/sort_ncu.log 2>&1
