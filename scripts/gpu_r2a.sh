# round 2, call 1: hygiene changes + the fused-aggregation epilogue variants (agg_ilp 1 / 2, CTA pairs + 16 epilogue warps)
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
B="python bench.py --steps 10 --no-cpu-baseline --no-e2e --no-projection"
i=0
for v in "FITGNN_AGG_ILP=1" "FITGNN_AGG_ILP=2" "FITGNN_AGG_WIDE=2 FITGNN_AGG_ILP=1" "FITGNN_AGG_ILP=2"; do
  i=$((i+1))
  echo "== variant $i: $v"
  env $v timeout 600 python -m pytest tests/test_gpu_aligned.py -x -q -k "transform_aggregate or fused_aggregation" 2>&1 | tail -1
  env $v timeout 600 $B > gpurun_out/bench_r2a_$i.log 2>&1; python scripts/show_bench.py gpurun_out/bench_r2a_$i.log | head -6
done
