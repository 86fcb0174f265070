timeout 600 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_aligned.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --no-cpu-baseline --no-projection --no-e2e > gpurun_out/bench_r1y.log 2>&1; python scripts/show_bench.py gpurun_out/bench_r1y.log 2>&1 | head -7
B="python bench.py --steps 2 --no-cpu-baseline --no-e2e --no-projection --profiler-range"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -c 8 --csv --log-file gpurun_out/dram_r1y.csv $B > gpurun_out/ncu_r1y.log 2>&1
