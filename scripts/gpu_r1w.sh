timeout 600 python -m pytest tests/test_gpu_aligned.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --no-cpu-baseline --no-projection --no-e2e > gpurun_out/bench_r1w.log 2>&1; python scripts/show_bench.py gpurun_out/bench_r1w.log 2>&1 | head -7
