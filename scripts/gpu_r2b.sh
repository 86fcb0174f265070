# round 2, call 2: streamed / hybrid engine, tensor-core defaults, bench mode blocks
timeout 900 python -m pytest tests/test_gpu_stream.py -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_stream.py 2>&1 | tail -15
timeout 600 python bench.py --workload products-small --steps 5 --cpu-seconds 3 > gpurun_out/bench_r2b_small.log 2>&1; tail -c 3000 gpurun_out/bench_r2b_small.log
timeout 900 python bench.py --steps 10 > gpurun_out/bench_r2b_full.log 2>&1; tail -c 6000 gpurun_out/bench_r2b_full.log
