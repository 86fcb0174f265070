# round 1, final code of session 3 (after the branch-free, pad-filling group-local aggregation): full GPU suite, default bench (both arms), ncu launch list + full capture
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1x_full.log 2>&1; python scripts/show_bench.py gpurun_out/bench_r1x_full.log | head -9
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1x_reference.log 2>&1; tail -1 gpurun_out/bench_r1x_reference.log | cut -c1-400
B="python bench.py --steps 2 --no-cpu-baseline --no-e2e --no-projection --profiler-range"
$B > gpurun_out/plain_r1x.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 200 --csv --log-file gpurun_out/launches_r1x.csv $B > gpurun_out/ncu_r1x_1.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -c 4 -o gpurun_out/fwd_r1x $B > gpurun_out/ncu_r1x_2.log 2>&1
tail -2 gpurun_out/ncu_r1x_2.log
