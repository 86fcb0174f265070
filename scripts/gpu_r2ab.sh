# round 2: 16-byte-load fp16 SpMM, 12-warp small-K f16 transform: tests + heavy-tail / cluster blocks
timeout 1200 python -m pytest tests/test_gpu_stream.py tests/test_gpu_gemm_tc.py tests/test_gpu_aligned.py -x -q 2>&1 | tail -5
timeout 900 python bench.py --steps 10 --only-modes --modes none_heavy_tail,cluster --mode-steps 5 > gpurun_out/bench_r2ab_modes.log 2>&1; python scripts/show_modes.py gpurun_out/bench_r2ab_modes.log
