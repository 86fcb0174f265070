# W-stationary pairs vs streaming pairs with one W plane: full-step A/B + ncu of both plans
FITGNN_GEMM_PAIR_WS=0 timeout 900 python bench.py --steps 10 --warmup 3 --modes= --no-cpu-baseline --no-projection > gpurun_out/bench_r2af_stream.log 2> gpurun_out/bench_r2af.err; tail -3 gpurun_out/bench_r2af.err
python - <<'PY'
import json
l = json.loads(open("gpurun_out/bench_r2af_stream.log").read().strip().splitlines()[-1])
print(l["ms_per_step"], l["value"], l["clocks"])
for k, v in l["kernels"].items(): print(k, round(v["ms"], 3), round(v["GBps"]), round(v["TFLOPs"], 1))
PY
NCU=1 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16x3_kernel -c 4 -o gpurun_out/gemm1_r2af python scripts/bench_gemm1_f16.py > gpurun_out/ncu_r2af.log 2>&1
tail -2 gpurun_out/ncu_r2af.log | cut -c1-200
