# round 2: block-dense tensor-core aggregation (spmm_mma)
timeout 900 python -m pytest tests/test_gpu_stream.py -x -q 2>&1 | tail -5
timeout 900 python bench.py --steps 10 --only-modes --modes cluster --mode-steps 3 > gpurun_out/bench_r2i_cluster.log 2>&1; python scripts/show_modes.py gpurun_out/bench_r2i_cluster.log
