timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/bench_r1m.log 2>&1; python scripts/show_bench.py gpurun_out/bench_r1m.log 2>&1 | head -7
