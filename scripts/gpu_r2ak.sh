# conv_fused schedule (aggregation on the accumulators of the hidden -> hidden pair kernel): suite + A/B; prefetch default still on here
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for cfg in "1 0" "0 0"; do
set -- $cfg
FITGNN_CONV_FUSED=$1 FITGNN_GEMM_PREFETCH=$2 timeout 900 python bench.py --steps 10 --warmup 3 --modes= --no-cpu-baseline --no-projection > gpurun_out/bench_r2ak_cf$1_pf$2.log 2> gpurun_out/bench_r2ak.err; tail -3 gpurun_out/bench_r2ak.err
python - <<PY
import json
l = json.loads(open("gpurun_out/bench_r2ak_cf$1_pf$2.log").read().strip().splitlines()[-1])
print("conv_fused=$1 prefetch=$2", round(l["ms_per_step"], 3), l["clocks"]["reasons"], " ".join(f"{k}={v['ms']:.3f}" for k, v in l["kernels"].items()))
PY
done
