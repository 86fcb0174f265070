# round 2, call 7: blocked SpMM v4 + training kernels + train bench block
timeout 900 python -m pytest tests/test_gpu_stream.py tests/test_gpu_training.py -x -q 2>&1 | tail -12
timeout 900 python bench.py --steps 10 --only-modes --modes cluster,train --mode-steps 3 > gpurun_out/bench_r2f_cluster.log 2>&1; python scripts/show_modes.py gpurun_out/bench_r2f_cluster.log
