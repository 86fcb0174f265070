# round 2 (session 2): single fp16 weight plane + W-stationary CTA pairs — parity tests, micro-benchmark, headline bench
timeout 900 python -m pytest tests/test_gpu_stream.py tests/test_gpu_gemm_tc.py -m gpu -x -q -k "fp16 or f16 or gemm" 2>&1 | tail -6
timeout 300 python scripts/bench_gemm1_f16.py 2>&1 | tail -6
timeout 900 python bench.py --steps 10 --warmup 3 --modes= > gpurun_out/bench_r2ae.log 2> gpurun_out/bench_r2ae.err; tail -3 gpurun_out/bench_r2ae.err
python - <<'PY'
import json
l = json.loads(open("gpurun_out/bench_r2ae.log").read().strip().splitlines()[-1])
print(l["ms_per_step"], l["value"], l["e2e"]["value"], l.get("parity"))
for k, v in l["kernels"].items(): print(k, round(v["ms"], 3), round(v["GBps"]), round(v["TFLOPs"], 1))
print(l["roofline"])
PY
