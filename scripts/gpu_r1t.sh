timeout 600 python -m pytest tests/test_gpu_aligned.py -m gpu -x -q 2>&1 | tail -6
timeout 900 python bench.py --steps 10 --no-cpu-baseline --no-projection > gpurun_out/bench_r1t.log 2>&1; python scripts/show_bench.py gpurun_out/bench_r1t.log 2>&1 | head -12
timeout 900 python bench.py --steps 10 --no-cpu-baseline --no-projection --no-e2e --features table > gpurun_out/bench_r1t_table.log 2>&1; python scripts/show_bench.py gpurun_out/bench_r1t_table.log 2>&1 | head -3
