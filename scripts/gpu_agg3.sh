timeout 900 python -m pytest tests/test_gpu_aligned.py tests/test_gpu_gemm_tc.py -x -q 2>&1 | tail -4
B="python bench.py --steps 10 --no-cpu-baseline --no-e2e --no-projection"
timeout 600 $B > gpurun_out/bench_agg3a.log 2>&1; python scripts/show_bench.py gpurun_out/bench_agg3a.log
FITGNN_AGG_WIDE=0 timeout 600 $B > gpurun_out/bench_agg3b.log 2>&1; python scripts/show_bench.py gpurun_out/bench_agg3b.log
FITGNN_GEMM_WIDE=1 timeout 600 $B --no-fuse-aggregate > gpurun_out/bench_agg3c.log 2>&1; python scripts/show_bench.py gpurun_out/bench_agg3c.log
timeout 600 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/bench_agg3d.log 2>&1; python scripts/show_bench.py gpurun_out/bench_agg3d.log
