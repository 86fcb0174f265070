# round 2, call 4: blocked SpMM v2 (8 lanes per row), cache reader, C forward
timeout 900 python -m pytest tests/test_gpu_stream.py -x -q 2>&1 | tail -8
timeout 900 python bench.py --steps 10 --only-modes --modes cluster --mode-steps 3 > gpurun_out/bench_r2d_cluster.log 2>&1; python scripts/show_modes.py gpurun_out/bench_r2d_cluster.log
