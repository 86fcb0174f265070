# round 2: dense aggregation A/B — tensor-core pieces (mma) vs shared-memory gathers (lds), compacted blocks
timeout 900 python -m pytest tests/test_gpu_stream.py -x -q 2>&1 | tail -4
for v in mma lds; do
  echo "== FITGNN_DENSE_SPMM=$v"
  FITGNN_DENSE_SPMM=$v timeout 900 python bench.py --steps 10 --only-modes --modes cluster --mode-steps 3 > gpurun_out/bench_r2k_cluster_$v.log 2>&1; python scripts/show_modes.py gpurun_out/bench_r2k_cluster_$v.log | head -8
done
