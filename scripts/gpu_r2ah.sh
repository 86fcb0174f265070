# incremental tile walk (no 64-bit divisions per tile) in the tcgen05 GEMM: whole GPU suite, micro-benchmark, headline
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python scripts/bench_gemm1_f16.py 2>&1 | tail -9
for ws in 1 0; do
FITGNN_GEMM_PAIR_WS=$ws timeout 900 python bench.py --steps 10 --warmup 3 --modes= --no-cpu-baseline --no-projection > gpurun_out/bench_r2ah_ws$ws.log 2> gpurun_out/bench_r2ah.err; tail -3 gpurun_out/bench_r2ah.err
python - <<PY
import json
l = json.loads(open("gpurun_out/bench_r2ah_ws$ws.log").read().strip().splitlines()[-1])
print("ws=$ws", l["ms_per_step"], l["value"], l["clocks"])
for k, v in l["kernels"].items(): print(k, round(v["ms"], 3), round(v["GBps"]), round(v["TFLOPs"], 1))
PY
done
