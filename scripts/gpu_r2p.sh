# round 2: fp16x2 as the bench headline — full default bench + peer test
timeout 600 python -m pytest tests/test_gpu_peer.py -x -q 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2p_full.log 2>&1; python scripts/show_modes.py gpurun_out/bench_r2p_full.log; tail -3 gpurun_out/bench_r2p_full.log | grep -v "^{" | cut -c1-500
