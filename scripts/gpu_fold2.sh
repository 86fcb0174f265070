timeout 900 python -m pytest tests/test_gpu_aligned.py -x -q 2>&1 | tail -3
for T in 1 0 1 0; do
FITGNN_ENGINE_TRICKS=$T timeout 900 python bench.py --steps 20 --no-cpu-baseline --no-projection --no-e2e > gpurun_out/bench_r1o_$T.log 2>&1; python scripts/show_bench.py gpurun_out/bench_r1o_$T.log 2>&1 | head -5
done
