# round 2: heavy-tail block with / without the block-staged SpMM for the classic part (3.5 entries per row)
for v in 8 3; do
  echo "== FITGNN_BLOCKED_MIN=$v"
  FITGNN_BLOCKED_MIN=$v timeout 900 python bench.py --steps 10 --only-modes --modes none_heavy_tail --mode-steps 5 > gpurun_out/bench_r2l_ht_$v.log 2>&1; python scripts/show_modes.py gpurun_out/bench_r2l_ht_$v.log | head -16
done
