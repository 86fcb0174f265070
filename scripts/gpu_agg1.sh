timeout 600 python -m pytest tests/test_gpu_aligned.py tests/test_gpu_gemm_tc.py -x -q 2>&1 | tail -25
timeout 900 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/bench_agg.log 2>&1; tail -c 600 gpurun_out/bench_agg.log
timeout 900 python bench.py --steps 10 --no-cpu-baseline --no-e2e --no-projection --no-fuse-aggregate > gpurun_out/bench_classic.log 2>&1; tail -c 300 gpurun_out/bench_classic.log
