timeout 300 python -m pytest tests/test_gpu_gemm_tc.py -x -q 2>&1 | tail -5
python bench.py --steps 10 --precision bf16x3 --no-cpu-baseline > gpurun_out/bench_tc3.log 2>&1; tail -c 300 gpurun_out/bench_tc3.log
