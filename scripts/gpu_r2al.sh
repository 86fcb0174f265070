# head with ONE weight plane (8 A stages instead of 5) vs hi/lo
for h in 1 0 1 0; do
FITGNN_HEAD_W1=$h timeout 900 python bench.py --steps 10 --warmup 3 --modes= --no-cpu-baseline --no-projection > gpurun_out/bench_r2al_h$h.log 2> gpurun_out/bench_r2al.err; tail -3 gpurun_out/bench_r2al.err
python - <<PY
import json
l = json.loads(open("gpurun_out/bench_r2al_h$h.log").read().strip().splitlines()[-1])
print("head_w1=$h", round(l["ms_per_step"], 3), l["clocks"]["reasons"], " ".join(f"{k}={v['ms']:.3f}" for k, v in l["kernels"].items()), l.get("parity"))
PY
done
