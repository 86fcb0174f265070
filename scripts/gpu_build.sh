timeout 300 python scripts/bench_build.py 2>&1 | tail -6
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/build_launches.csv python scripts/bench_build.py > gpurun_out/build_ncu.log 2>&1
