# round 2: full validation — GPU test suite, smoke, default bench, reference arm, ncu launch list + full capture
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2as_full.log 2> gpurun_out/bench_r2as_full.err; tail -2 gpurun_out/bench_r2as_full.err; python scripts/show_modes.py gpurun_out/bench_r2as_full.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2as_ref.log 2>&1; tail -1 gpurun_out/bench_r2as_ref.log | cut -c1-400
B="python bench.py --steps 2 --no-cpu-baseline --no-e2e --no-projection --modes= --profiler-range"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 200 --csv --log-file gpurun_out/launches_r2as.csv $B > gpurun_out/ncu_r2as_1.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -c 4 -o gpurun_out/fwd_r2as $B > gpurun_out/ncu_r2as_2.log 2>&1
tail -1 gpurun_out/ncu_r2as_2.log | cut -c1-200
