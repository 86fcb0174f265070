timeout 600 ncu --set full --clock-control none --import-source on -k regex:rs_ -s 2 -c 2 -o gpurun_out/sort_r1s python scripts/bench_sort.py 1.23e8 > gpurun_out/sort_ncu_full.log 2>&1
ls -la gpurun_out/ | tail -3
