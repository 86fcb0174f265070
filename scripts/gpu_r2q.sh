# round 2: ncu launch list + full capture of the fp16x2 headline; heavy-tail block re-run (hybrid with fp16x2)
B="python bench.py --steps 2 --no-cpu-baseline --no-e2e --no-projection --modes= --profiler-range"
$B > gpurun_out/plain_r2q.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 200 --csv --log-file gpurun_out/launches_r2q.csv $B > gpurun_out/ncu_r2q_1.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -c 4 -o gpurun_out/fwd_r2q $B > gpurun_out/ncu_r2q_2.log 2>&1
tail -2 gpurun_out/ncu_r2q_2.log
timeout 600 python bench.py --steps 10 --only-modes --modes none_heavy_tail --mode-steps 5 > gpurun_out/bench_r2q_ht.log 2>&1; python scripts/show_modes.py gpurun_out/bench_r2q_ht.log
