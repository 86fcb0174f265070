# round 2: fp16 hidden state (opt-in) — tests + bench block
timeout 900 python -m pytest tests/test_gpu_stream.py tests/test_gpu_gemm_tc.py tests/test_gpu_aligned.py -x -q 2>&1 | tail -12
timeout 900 python bench.py --steps 10 --only-modes --modes fp16x2 --mode-steps 10 > gpurun_out/bench_r2o_fp16x2.log 2>&1; python scripts/show_modes.py gpurun_out/bench_r2o_fp16x2.log; tail -3 gpurun_out/bench_r2o_fp16x2.log | cut -c1-600
