# pair kernel: 8 epilogue warps + 5 A stages vs 12 warps + 4 stages (W-stationary), after the walk / issue-loop fixes
for w in -1 0 -1 0; do
FITGNN_GEMM_WIDE=$w timeout 900 python bench.py --steps 10 --warmup 3 --modes= --no-cpu-baseline --no-projection > gpurun_out/bench_r2an_w$w.log 2> gpurun_out/bench_r2an.err; tail -3 gpurun_out/bench_r2an.err
python - <<PY
import json
l = json.loads(open("gpurun_out/bench_r2an_w$w.log").read().strip().splitlines()[-1])
print("gemm_wide=$w", round(l["ms_per_step"], 3), l["clocks"]["reasons"], " ".join(f"{k}={v['ms']:.3f}" for k, v in l["kernels"].items()))
PY
done
