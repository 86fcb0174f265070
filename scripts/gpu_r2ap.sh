# exchange loop of the 12-warp aggregation epilogue in 16-column chunks (default build) vs 32-column chunks (_w32 library)
for lib in "" "_w32" "" "_w32"; do
FITGNN_B200_LIB=$PWD/fitgnn_b200/libfitgnn_b200$lib.so timeout 900 python bench.py --steps 10 --warmup 3 --modes= --no-cpu-baseline --no-projection > gpurun_out/bench_r2ap$lib.log 2> gpurun_out/bench_r2ap.err; tail -3 gpurun_out/bench_r2ap.err
python - <<PY
import json
l = json.loads(open("gpurun_out/bench_r2ap$lib.log").read().strip().splitlines()[-1])
print("lib='$lib'", round(l["ms_per_step"], 3), l["clocks"]["reasons"], " ".join(f"{k}={v['ms']:.3f}" for k, v in l["kernels"].items()))
PY
done
