# round 2: full validation — GPU test suite, smoke, default bench (both arms), ncu launch list + full captures
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2m_full.log 2>&1; python scripts/show_modes.py gpurun_out/bench_r2m_full.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2m_reference.log 2>&1; tail -1 gpurun_out/bench_r2m_reference.log | cut -c1-300
B="python bench.py --steps 2 --no-cpu-baseline --no-e2e --no-projection --modes= --profiler-range"
$B > gpurun_out/plain_r2m.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 200 --csv --log-file gpurun_out/launches_r2m.csv $B > gpurun_out/ncu_r2m_1.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -c 4 -o gpurun_out/fwd_r2m $B > gpurun_out/ncu_r2m_2.log 2>&1
tail -2 gpurun_out/ncu_r2m_2.log
