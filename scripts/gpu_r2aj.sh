# L2 prefetch in the TMA producer, specialised MMA issue loop, bias loads under the TMEM-load wait: suite + A/B
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python scripts/bench_gemm1_f16.py 2>&1 | grep "12 epi"
for cfg in "1 1" "0 1" "1 0" "0 0"; do
set -- $cfg
FITGNN_GEMM_PAIR_WS=$1 FITGNN_GEMM_PREFETCH=$2 timeout 900 python bench.py --steps 10 --warmup 3 --modes= --no-cpu-baseline --no-projection > gpurun_out/bench_r2aj_ws$1_pf$2.log 2> gpurun_out/bench_r2aj.err; tail -3 gpurun_out/bench_r2aj.err
python - <<PY
import json
l = json.loads(open("gpurun_out/bench_r2aj_ws$1_pf$2.log").read().strip().splitlines()[-1])
print("ws=$1 prefetch=$2", round(l["ms_per_step"], 3), l["clocks"]["reasons"], " ".join(f"{k}={v['ms']:.3f}" for k, v in l["kernels"].items()))
PY
done
