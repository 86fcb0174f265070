B="python bench.py --steps 2 --no-cpu-baseline --no-e2e --no-projection --profiler-range"
$B > gpurun_out/plain_r1g.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 200 --csv --log-file gpurun_out/launches_r1g.csv $B > gpurun_out/ncu_r1g_1.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -c 4 -o gpurun_out/fwd_r1g $B > gpurun_out/ncu_r1g_2.log 2>&1
tail -2 gpurun_out/ncu_r1g_2.log
python bench.py --steps 10 > gpurun_out/bench_r1g_full.log 2>&1; python scripts/show_bench.py gpurun_out/bench_r1g_full.log | head -8
