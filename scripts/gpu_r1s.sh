# round-1 session 3: new radix sort + compact Ac keys + pack-ordered features
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 900 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/bench_r1s.log 2>&1; python scripts/show_bench.py gpurun_out/bench_r1s.log 2>&1 | head -12
