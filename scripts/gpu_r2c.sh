# round 2, call 3: blocked SpMM, one-call C forward, hybrid fix, graph-level training, load_graph_data
timeout 900 python -m pytest tests/test_gpu_stream.py tests/test_gpu_training.py -x -q 2>&1 | tail -15
timeout 900 python bench.py --steps 10 --no-e2e --no-projection --cpu-seconds 5 > gpurun_out/bench_r2c_full.log 2>&1; python scripts/show_modes.py gpurun_out/bench_r2c_full.log
