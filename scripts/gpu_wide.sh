for W in 0 1; do
FITGNN_AGG_WIDE=$W python bench.py --steps 10 --no-cpu-baseline --no-projection --no-e2e > gpurun_out/bench_r1k_wide$W.log 2>&1; python scripts/show_bench.py gpurun_out/bench_r1k_wide$W.log 2>&1 | head -5
done
