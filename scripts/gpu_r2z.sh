# round 2: fp16 SpMM + classic fp16 schedule tests; heavy-tail / cluster blocks with it; peer test (cluster-launched push)
timeout 1200 python -m pytest tests/test_gpu_stream.py tests/test_gpu_parity.py tests/test_gpu_peer.py -x -q 2>&1 | tail -5
timeout 900 python bench.py --steps 10 --only-modes --modes none_heavy_tail,cluster --mode-steps 5 > gpurun_out/bench_r2z_modes.log 2>&1; python scripts/show_modes.py gpurun_out/bench_r2z_modes.log
