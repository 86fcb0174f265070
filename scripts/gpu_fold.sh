timeout 900 python -m pytest tests/test_gpu_aligned.py tests/test_gpu_peer.py tests/test_gpu_gemm_tc.py -x -q 2>&1 | tail -5
timeout 900 python bench.py --steps 10 --no-cpu-baseline --no-projection > gpurun_out/bench_r1n.log 2>&1; python scripts/show_bench.py gpurun_out/bench_r1n.log 2>&1 | head -7
